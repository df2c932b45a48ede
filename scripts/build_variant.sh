#!/bin/bash
# Build a variant of libpf_cuda.so into build/variants/ (development tool for scripts/exp_build_variants.py).
#   scripts/build_variant.sh NAME [GIT_REV|WORK] [EXTRA nvcc flags...]
set -e
name=$1; rev=${2:-WORK}; shift; shift || true
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p "$tmp/simpeg_b200" "$tmp/include" "$root/build/variants"
if [ "$rev" = WORK ]; then
    cp -r "$root/simpeg_b200/csrc" "$tmp/simpeg_b200/csrc"; cp "$root"/include/*.h "$tmp/include/"
else
    git -C "$root" archive "$rev" simpeg_b200/csrc include | tar -x -C "$tmp"
fi
rm -f "$tmp"/simpeg_b200/csrc/*.o
make -C "$tmp/simpeg_b200/csrc" -j5 OUT="$root/build/variants/libpf_$name.so" EXTRA="$*" 2>&1 | grep -E "error|warning: v" || true
rm -rf "$tmp"
ls -la "$root/build/variants/libpf_$name.so"
