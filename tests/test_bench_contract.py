"""CPU: bench.py's reference arm runs without a GPU and prints one JSON line with the
contract's keys (the CUDA arm is exercised on the GPU box by the driver)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_json_line():
    out = subprocess.run(
        [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "cfg1",
         "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 1e6
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run(
        [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "cfg1",
         "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_clock_sampler_selects_the_timed_region(tmp_path):
    """bench.ClockSampler: samples are picked by timestamp inside the timed region; when the region
    caught none (8-GPU runs: nvidia-smi comes up slower than the region lasts) the warm-up + timed
    samples are used and labelled; throttle reasons and the under-load median follow the kept rows."""
    import datetime
    import sys as _sys

    _sys.path.insert(0, str(ROOT))
    import bench

    class Done:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    t0 = datetime.datetime(2026, 1, 1, 12, 0, 0)
    rows = [
        (t0 + datetime.timedelta(milliseconds=100), 1200, 1965, 150.0, "Not Active", "Not Active", "Not Active", "Not Active"),
        (t0 + datetime.timedelta(milliseconds=600), 1950, 1965, 510.0, "Not Active", "Not Active", "Not Active", "Active"),
        (t0 + datetime.timedelta(milliseconds=700), 1965, 1965, 530.0, "Not Active", "Not Active", "Not Active", "Not Active"),
        (t0 + datetime.timedelta(milliseconds=1500), 1400, 1965, 200.0, "Active", "Not Active", "Not Active", "Not Active"),
    ]

    def sampler():
        path = tmp_path / "smi.csv"
        with open(path, "w") as fh:
            fh.write("garbage line\n")
            for r in rows:
                fh.write(", ".join([r[0].strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]] + [str(v) for v in r[1:]]) + "\n")
        s = bench.ClockSampler(0)
        s.proc, s.path = Done(), str(path)
        return s

    inside = sampler().stop((t0 + datetime.timedelta(milliseconds=500), t0 + datetime.timedelta(milliseconds=800)))
    assert inside["window"] == "timed region" and inside["samples"] == 2
    assert inside["sm_mhz"] == 1957.5 and inside["sm_max_mhz"] == 1965.0 and inside["reasons"] == ["sw_power_cap"]
    none = sampler().stop((t0 + datetime.timedelta(seconds=5), t0 + datetime.timedelta(seconds=6)))
    assert none["window"] == "warm-up + timed region" and none["samples"] == 4
    assert none["sm_mhz"] == 1957.5                      # idle samples (< half of the peak power) do not count
    assert none["reasons"] == ["hw_slowdown", "sw_power_cap"]
    empty = bench.ClockSampler(0)
    empty.proc, empty.path = Done(), str(tmp_path / "empty.csv")
    open(empty.path, "w").close()
    assert empty.stop((t0, t0))["reasons"] == ["no samples"]
