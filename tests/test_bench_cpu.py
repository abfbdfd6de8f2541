"""CPU suite for bench.py's contract: the reference arm (the CPU port of the reference's retrieval path, oracle/fast_cpu.py)
runs here without a GPU; its JSON line must carry the keys the driver reads and the SAME `config` object our arm prints
for the same arguments (the driver compares the two arms' configs), and the other ranks of a torchrun launch exit 0
without printing.  Our own arm must refuse to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--rows", "30000", "--dim", "96", "--k", "7", "--steps", "2", "--warmup", "1"]


def _run(extra, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + ARGS + extra, capture_output=True, text=True,
                          timeout=300, env=e, cwd=ROOT)


def _bench_module():
    sys.path.insert(0, ROOT)
    try:
        import bench
    finally:
        sys.path.pop(0)
    return bench


def test_reference_arm_line_and_shared_config():
    r = _run(["--impl", "reference"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step"] + 1e-9
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["extrapolated"] is False
    assert cb["rows_timed"] == 30000 and "30000-row corpus" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    bench = _bench_module()
    a = types.SimpleNamespace(rows=30000, dim=96, dtype="bf16", batch=1, k=7)
    assert d["config"] == bench.job_config(a, 1)                  # what our arm prints under "config" for these arguments
    assert d["metric"] == bench.METRIC
    assert set(d["config"]) == {"workload", "rows", "dim", "k", "batch", "parallelism", "l2"}


def test_reference_arm_runs_on_rank_0_only():
    r = _run(["--impl", "reference", "--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
    r = _run(["--impl", "reference", "--gpus", "2"], env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"})
    assert r.returncode == 0
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["n_gpus"] == 2 and d["config"]["parallelism"] == "row-shard x2"


def test_our_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return                                                     # the GPU suite runs the real thing
    r = _run(["--no-cpu-baseline", "--no-extra-regimes", "--also-batch", "0"])
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")], "a bench line was printed without a GPU"
