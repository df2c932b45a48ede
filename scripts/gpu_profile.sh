#!/bin/bash
# Profiling recipe of the round (run under gpurun on ONE GPU); outputs land in gpurun_out/.
set -x
TAG=${1:-r1}
# 1) launch list of the bench command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# 2) full capture of the dominant kernels (build, gemv, gemv_t)
ncu --set full --clock-control none --import-source on -k regex:"pf_march_kernel|pf_tile_kernel|pf_node_gtv_kernel" -c 3 \
    -o gpurun_out/prof_build_${TAG} python scripts/prof_build.py 296 > gpurun_out/prof_build_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemv_kernel|column_kernel' -c 2 \
    -o gpurun_out/prof_products_${TAG} python scripts/prof_build.py 296 > gpurun_out/prof_products_${TAG}.log 2>&1
# 3) the bench itself (not under a profiler) + the reference arm
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.json
# 4) fp32-stored G products and the draped equivalent-source layer (tile kernel)
ncu --set full --clock-control none --import-source on -k regex:'gemv_kernel|column_kernel' -c 2 \
    -o gpurun_out/prof_f32_${TAG} python scripts/prof_f32.py > gpurun_out/prof_f32_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pf_tile_kernel' -c 1 \
    -o gpurun_out/prof_layer_${TAG} python scripts/prof_layer.py > gpurun_out/prof_layer_${TAG}.log 2>&1
