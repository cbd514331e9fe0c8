"""bench.py's contract as the driver uses it, as far as it can be checked without a GPU: the reference arm
(`--impl reference`) runs the oracle's kernels on the host cores and prints ONE JSON line with the keys the driver reads;
the flop / task-count helpers agree with the expanded DAG."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")          # what torchrun sets: the arm must not depend on it
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-n", "8192"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "TFLOP/s"
    assert d["metric"].startswith("fp64 TFLOP/s Cholesky N=131072 tile=4096")
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "N=131072" in d["config"]["workload"] and d["config"]["extrapolated"] is True
    assert d["vs_baseline"] is None


def test_flop_and_task_counts_match_the_dag():
    import bench
    from numpywren_b200 import alg_wrappers
    from numpywren_b200.matrix import BigMatrix
    nb = 8
    A = BigMatrix("bench_contract_A", shape=(nb * 4, nb * 4), shard_sizes=(4, 4), device="cpu")
    program, _ = alg_wrappers.cholesky(A)
    names = [n.call.compute_name for n in program.program.nodes]
    c, r, s = bench.chol_task_counts(nb)
    assert (c, r, s) == (names.count("chol"), names.count("trsm"), names.count("syrk"))
    assert bench.chol_flops(4096) == pytest.approx(4096 ** 3 / 3.0, rel=1e-3)
