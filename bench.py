#!/usr/bin/env python
"""bench.py — headline benchmark: fp64 TFLOP/s of the LambdaPACK blocked Cholesky on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cholesky|tsqr|gemm]
                  [--size N] [--tile B]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one complete factorisation of a synthetic SPD matrix A = X X^T + N I through the
reference-facing surface: alg_wrappers.cholesky(A) -> program.start() -> job_runner.lambdapack_run(program).
The workload is the SAME at every GPU count — BASELINE.json's metric: N=131072, tile 4096 (70.9 GB of lower tiles: it
fits one B200) — so the 1/2/4/8-GPU values form a strong-scaling curve.  TFLOP/s are algorithmic: N^3/3 flops per
factorisation divided by device time (CUDA events, barrier on both sides, max over ranks).

The JSON line also carries:
  config.parity_vs_oracle — ||L_gpu - L_oracle||_F / ||L_oracle||_F for a whole factorisation at the benchmark tile
                 (N=16384, tile 4096) on the same GPUs / process grid as the timed run, against the CPU oracle's factor
                 of the same host tiles (must be <= 1e-10); plus a size-independent residual at the full size;
  roofline     — the dominant kernel (gemm_nt_tma_kernel behind kernels.syrk, 2*b^3 flops per launch) timed alone with
                 CUDA events on its stream, against the fp64 tensor-pipe peak measured live (DMMA issue probe and cuBLAS
                 DGEMM; MEASURED_PEAKS.json has no fp64 entry; the datasheet fraction is printed beside it);
  cpu_baseline — the CPU oracle's kernels (NumPy/SciPy restatement of the reference path, pinned to the reference's
                 outputs) timed on all host cores at the benchmark tile and composed by the workload's task counts
                 ("extrapolate by task counts; state the extrapolation"), with a measured N=16384 end-to-end
                 oracle run as cross-check;
  e2e          — the same factorisation with HOST (pinned) input/output tiles, host<->device copies inside the timed
                 region, at this GPU count (every rank stages its own share).

--impl reference times the CPU arm alone, on the same config/metric (rank 0 only under torchrun).
--workload tsqr / gemm run BASELINE configs 4 / 5 (TSQR 4194304 x 512 tile (65536, 512); GEMM tile 8192) the same way.
"""
from __future__ import annotations

import os
import sys

if "reference" in sys.argv:
    # the CPU arm uses every host core, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse  # noqa: E402
import contextlib  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
T_START = time.time()

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "fp64 TFLOP/s Cholesky N=131072 tile=4096 at 1/2/4/8 B200; % of fp64 peak"
UNIT = "TFLOP/s"
FP64_DATASHEET_TFLOPS = 37.0          # HGX B200 datasheet, fp64 / fp64 tensor core, per GPU (VERDICT r1 quotes it)
PARITY_BAR = 1e-10


# ----------------------------------------------------------------------------------------------- helpers
def _dtype_label():
    """Arithmetic the path computes in.  "f64" unless the int8-tensor-core emulation of the syrk products was switched on
    explicitly (NPW_B200_SYRK=i8emu) — then the line says so instead of claiming plain fp64."""
    if os.environ.get("NPW_B200_SYRK", "native") == "i8emu":
        return "f64 emulated on int8 tensor cores for the syrk products (%s digits; trsm/potrf native f64)" % os.environ.get(
            "NPW_B200_I8_DIGITS", "8")
    return "f64"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def budget_left():
    """Seconds left of the wall-clock budget this process gives itself (the driver's per-run limit is ~870 s)."""
    return float(os.environ.get("NPW_B200_BENCH_BUDGET_S", "780")) - (time.time() - T_START)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0, period_ms=200):
        self.index, self.period_ms = index, period_ms
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # under load = the upper half of the samples (idle samples at the edges are dropped)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


def chol_flops(n):
    return n ** 3 / 3.0


def chol_task_counts(nb):
    """(chol, trsm, syrk) tile tasks of algs.CHOLESKY on an nb x nb tile grid."""
    return nb, nb * (nb - 1) // 2, (nb - 1) * nb * (nb + 1) // 6


# ----------------------------------------------------------------------------------------------- distributed context
class Ctx:
    """One process per GPU.  world == 1: plain single-GPU run, no process group."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
        self.device = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.device)
        self.grid = None
        if self.world > 1:
            from numpywren_b200 import parallel
            self.grid = parallel.init_from_env("nccl")
        from numpywren_b200 import _capi
        _capi.load()                     # fails loudly when libnpw_b200.so is missing

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(device_ids=[self.device.index])

    def max_float(self, v):
        if self.world == 1:
            return float(v)
        import torch.distributed as dist
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def min_int(self, v):
        if self.world == 1:
            return int(v)
        import torch.distributed as dist
        t = torch.tensor([int(v)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    def sum_int(self, v):
        if self.world == 1:
            return int(v)
        import torch.distributed as dist
        t = torch.tensor([int(v)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def bcast_tile(self, t, src):
        if self.world > 1:
            import torch.distributed as dist
            dist.broadcast(t, src=src)
        return t

    def is_mine(self, m, idx):
        return self.grid is None or self.grid.is_mine(m, idx)

    def owner(self, m, idx):
        return 0 if self.grid is None else self.grid.owner(m, idx)

    def finish(self):
        if self.world > 1:
            import torch.distributed as dist
            self.barrier()
            dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- CPU arm (oracle)
@contextlib.contextmanager
def all_host_cores():
    """BLAS on every host core for the CPU arm, whatever OMP_NUM_THREADS the launcher exported."""
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:  # pragma: no cover
        yield
        return
    with threadpool_limits(limits=os.cpu_count() or 1):
        yield


def threads_in_use():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))


def spd_panels(nb, b, width=128):
    """Row blocks X_j of the benchmark generator A = X X^T + n I, X_j = RandomState(j).randn(b, width) (the reference
    tests' recipe, tests/test_failures.py:33-37, per-row-block seeds: SURVEY §8d)."""
    return [np.random.RandomState(j).randn(b, width) for j in range(nb)]


def spd_host_tile(X, j, k, n):
    t = X[j].dot(X[k].T)
    if j == k:
        t[np.diag_indices(t.shape[0])] += n
    return t


def cpu_cholesky_sample(n, b):
    """The CPU oracle's whole Cholesky (reference kernels in program order) on an n x n sample with tile b, on all host
    cores.  -> (seconds, {(j,k): factor tile}, {(j,k): input tile})."""
    from oracle import npw_oracle as orc
    nb = n // b
    X = spd_panels(nb, b)
    I = orc.OracleBigMatrix("I", (n, n), (b, b))
    tiles = {}
    for j in range(nb):
        for k in range(j + 1):
            tiles[(j, k)] = spd_host_tile(X, j, k, n)
            I.store[(j, k)] = tiles[(j, k)].copy()
    with all_host_cores():
        t0 = time.perf_counter()
        O, _ = orc.run_cholesky(I)
        dt = time.perf_counter() - t0
        threads = threads_in_use()
    L = {(j, k): np.ascontiguousarray(O.get_block(j, k)) for j in range(nb) for k in range(j + 1)}
    return dt, L, tiles, threads


class CpuKernelSample:
    """One sample = the oracle's syrk, trsm and chol (reference kernels.py:212-257) once each on b x b tiles, all host
    cores.  The whole-workload CPU time is composed from these by the task counts of the DAG."""

    def __init__(self, b):
        from oracle import npw_oracle as orc
        self.orc, self.b = orc, b
        X = spd_panels(2, b)
        self.a00 = spd_host_tile(X, 0, 0, 2 * b)
        self.a10 = spd_host_tile(X, 1, 0, 2 * b)
        self.a11 = spd_host_tile(X, 1, 1, 2 * b)
        with all_host_cores():
            self.l00 = orc.chol(self.a00)
            self.l10 = np.ascontiguousarray(orc.trsm(self.l00, self.a10))
            self.threads = threads_in_use()

    def once(self):
        orc = self.orc
        with all_host_cores():
            t0 = time.perf_counter(); orc.syrk(self.a11, self.l10, self.l10); t1 = time.perf_counter()
            orc.trsm(self.l00, self.a10); t2 = time.perf_counter()
            orc.chol(self.a00); t3 = time.perf_counter()
        return {"syrk": t1 - t0, "trsm": t2 - t1, "chol": t3 - t2}

    @staticmethod
    def compose(t, nb):
        c, r, s = chol_task_counts(nb)
        return c * t["chol"] + r * t["trsm"] + s * t["syrk"]


def cpu_config1_gemm():
    """BASELINE config 1: 4096 x 4096 fp64 GEMM, BigMatrix tile 1024, the legacy binops.gemm(local=True) path
    (reference binops.py:19-33,155-158) as restated by the oracle, all host cores."""
    from oracle import npw_oracle as orc
    n, b = 4096, 1024
    rs = np.random.RandomState(0)
    A, B = rs.randn(n, n), rs.randn(n, n)
    X = orc.OracleBigMatrix("X", (n, n), (b, b)); orc.shard_matrix(X, A)
    Y = orc.OracleBigMatrix("Y", (n, n), (b, b)); orc.shard_matrix(Y, B)
    with all_host_cores():
        orc.binops_gemm(X, Y)
        t0 = time.perf_counter()
        orc.binops_gemm(X, Y)
        dt = time.perf_counter() - t0
    return {"workload": "GEMM 4096x4096 fp64, BigMatrix tile 1024 (binops.gemm local path)", "seconds": dt,
            "tflops": 2.0 * n ** 3 / dt * 1e-12}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; /root/reference cannot travel) on
    the SAME config as the GPU arm.  Each step is a bounded sample of that workload: one syrk, one trsm and one chol at
    the benchmark tile on all host cores; the step time is the workload's task counts x those kernel times."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.workload != "cholesky":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU arm is implemented for the headline Cholesky workload"}))
        return 0
    n, b = args.n, args.tile
    nb = n // b
    ks = CpuKernelSample(b)
    for _ in range(args.warmup):
        ks.once()
    samples = [ks.once() for _ in range(args.steps)]
    mean = {k: float(np.mean([s[k] for s in samples])) for k in ("syrk", "trsm", "chol")}
    dt = CpuKernelSample.compose(mean, nb)
    val = chol_flops(n) / dt * 1e-12
    cores = ks.threads
    c, r, s = chol_task_counts(nb)
    cross = None
    if budget_left() > 200:
        t_meas, _, _, _ = cpu_cholesky_sample(args.cpu_n, b)
        cross = {"measured_end_to_end_s": t_meas, "composed_s": CpuKernelSample.compose(mean, args.cpu_n // b),
                 "what": f"oracle run_cholesky N={args.cpu_n} tile={b}, measured whole vs composed from the same kernel times"}
    sample = (f"per step: oracle syrk + trsm + chol once each at {b}x{b} on {cores} host threads; step time = "
              f"{s} x syrk + {r} x trsm + {c} x chol (task counts of N={n}); EXTRAPOLATED, the full run is "
              f"{chol_flops(n):.2e} flop = hours on the host")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "kernel_seconds": mean, "extrapolated": True, "cross_check": cross},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm: shared pieces
def workload_name(args):
    if args.workload == "tsqr":
        return (f"TSQR {args.n}x{args.cols} fp64 (randn per tile), tile=({args.tile},{args.cols}), algs.TSQR LambdaPACK DAG "
                f"(binary reduction tree)")
    if args.workload == "gemm":
        return f"GEMM N={args.n} fp64 (randn per tile), tile={args.tile}, binops.gemm owner-computes / K-loop schedule (algs.GEMM_ACC)"
    return f"Cholesky N={args.n} fp64 SPD (A = X X^T + N I, X N x 128), tile={args.tile}, algs.CHOLESKY LambdaPACK DAG"


def measure_peaks(device):
    """fp64 roofline denominators measured live: DMMA issue probe (our C-ABI) and cuBLAS DGEMM 8192^3 (torch.matmul)."""
    import ctypes
    from numpywren_b200 import _capi
    lib = _capi.load()
    warps = 8
    scratch = torch.empty(max(1, lib.npw_fp64_pipe_probe_bytes(warps) // 8), dtype=torch.float64, device=device)
    fl = ctypes.c_double(0.0)
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _capi.check(lib.npw_fp64_pipe_probe(scratch.data_ptr(), 8192, warps, ctypes.byref(fl), torch.cuda.current_stream().cuda_stream), "probe")
        e1.record(); e1.synchronize()
        best = max(best, fl.value / e0.elapsed_time(e1) * 1e-9)
    a = torch.randn(8192, 8192, dtype=torch.float64, device=device)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=device)
    cb = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        cb = max(cb, 2.0 * 8192 ** 3 / e0.elapsed_time(e1) * 1e-9)
    del a, b
    return {"dmma_pipe_tflops": best, "cublas_dgemm_tflops": cb}


def measure_dominant_kernel(b, device, reps=12):
    """gemm_nt_tma_kernel (kernels.syrk on b x b tiles) timed alone on its stream; L2 is defeated by rotating over
    operand sets larger than the 126 MB L2 (3 x 128 MiB per launch, 4 sets)."""
    from numpywren_b200 import kernels as k
    sets = []
    for i in range(4):
        s = torch.empty(b, b, dtype=torch.float64, device=device); k.fill_random(s, 11 + i)
        x = torch.empty(b, b, dtype=torch.float64, device=device); k.fill_random(x, 21 + i)
        y = torch.empty(b, b, dtype=torch.float64, device=device); k.fill_random(y, 31 + i)
        sets.append((s, x, y))
    for s, x, y in sets[:3]:
        k.syrk(s, x, y, out=s)
    torch.cuda.synchronize()
    ts = []
    for r in range(reps):
        s, x, y = sets[r % 4]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); k.syrk(s, x, y, out=s); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts)), float(np.min(ts))


def load_traffic():
    p = os.path.join(ROOT, "profiles", "dominant_kernel_ncu.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def tensor_roofline(ctx, b, value):
    """Dominant kernel alone (rank 0's GPU) against the live-measured fp64 pipe peak."""
    peaks = measure_peaks(ctx.device)
    log("peaks", peaks)
    k_avg, k_min = measure_dominant_kernel(b, ctx.device)
    k_flops = 2.0 * b ** 3
    achieved = k_flops / (k_avg * 1e-3) * 1e-12
    peak = max(peaks["dmma_pipe_tflops"], peaks["cublas_dgemm_tflops"])
    return {"bound": "tensor", "kernel": "gemm_nt_tma_kernel (kernels.syrk, %d^3 tile update)" % b, "achieved": achieved,
            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": load_traffic(),
            "flops_per_launch": k_flops, "avg_launch_ms": k_avg, "min_launch_ms": k_min,
            "peak_source": "measured live on rank 0: max(DMMA.8x8x4 issue probe %.2f, cuBLAS DGEMM 8192^3 %.2f) TFLOP/s per GPU; "
                           "MEASURED_PEAKS.json has no fp64 entry" % (peaks["dmma_pipe_tflops"], peaks["cublas_dgemm_tflops"]),
            "frac_of_datasheet_%.0f_tflops" % FP64_DATASHEET_TFLOPS: achieved / FP64_DATASHEET_TFLOPS,
            "whole_job_frac_of_aggregate_peak": value / (peak * ctx.world),
            "whole_job_frac_of_aggregate_datasheet": value / (FP64_DATASHEET_TFLOPS * ctx.world)}


def free_all(*mats):
    for m in mats:
        m.free()


def host_pinned_cap_bytes(ctx):
    """Pinned host memory this rank may use for e2e staging: a fraction of what is available now, per local rank."""
    avail = 64 << 30
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
    except Exception:
        pass
    frac = float(os.environ.get("NPW_B200_BENCH_PINNED_FRAC", "0.6"))
    return int(avail * frac / max(1, ctx.world))


# ----------------------------------------------------------------------------------------------- Cholesky workload
class CholeskyWorkload:
    """Synthetic SPD input generated per tile on the owning GPU: A_jk = X_j X_k^T (+ n I on the diagonal)."""

    def __init__(self, ctx, n, b):
        from numpywren_b200 import kernels
        self.ctx, self.n, self.b, self.nb = ctx, n, b, n // b
        self.kernels = kernels
        self.X = [torch.empty(b, 128, dtype=torch.float64, device=ctx.device) for _ in range(self.nb)]
        for j in range(self.nb):
            kernels.fill_random(self.X[j], seed=20261017, row0=j * b)
        self.step_id = 0

    def tile(self, j, k):
        t = torch.empty(self.b, self.b, dtype=torch.float64, device=self.ctx.device)
        self.kernels._gemm_into(t, None, self.X[j], self.X[k], False, True, 1.0, 0.0)
        if j == k:
            self.kernels.add_diag(t, float(self.n))
        return t

    def new_matrix(self, tag):
        from numpywren_b200.matrix import BigMatrix
        self.step_id += 1
        return BigMatrix(f"bench_{tag}_{self.step_id}", shape=(self.n, self.n), shard_sizes=(self.b, self.b), device=self.ctx.device)

    def lower(self):
        return [(j, k) for j in range(self.nb) for k in range(j + 1)]

    def resident_input(self):
        """This rank's lower tiles of A generated directly in HBM (the timed region starts with inputs resident)."""
        A = self.new_matrix("A")
        for (j, k) in self.lower():
            if self.ctx.is_mine(A, (j, k)):
                A._put_block_ref(self.tile(j, k), j, k)
        return A


def run_program(ctx, program, streams, consume=True, profile=False):
    """prepare (static DAG analysis, outside the timed region) -> barrier -> [start, lambdapack_run] timed -> max over ranks."""
    from numpywren_b200 import _capi, job_runner
    from numpywren_b200 import lambdapack as lp
    _ = program.program.nodes
    plan_s = job_runner.prepare(program, streams=streams, consume_inputs=consume, profile=profile)
    torch.cuda.synchronize()
    ctx.barrier()
    l0 = _capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    program.start()
    job_runner.lambdapack_run(program, timeout=3600, streams=streams, consume_inputs=consume, profile=profile)
    e1.record()
    e1.synchronize()
    ms = ctx.max_float(e0.elapsed_time(e1))
    launches = ctx.sum_int(_capi.launch_count() - l0)
    assert program.program_status() == lp.PS.SUCCESS
    return ms, launches, plan_s


def cholesky_step(ctx, wl, streams):
    from numpywren_b200.alg_wrappers import cholesky
    A = wl.resident_input()
    program, meta = cholesky(A)
    ms, launches, plan_s = run_program(ctx, program, streams)
    return ms, launches, plan_s, A, program, meta


def residual_check(ctx, wl, O):
    """||(L L^T)_jj - A_jj|| / ||A_jj|| on the LAST diagonal tile (it depends on every panel of the factor): the
    size-independent parity property at the full benchmark size.  Row tiles are gathered from their owners."""
    j = wl.nb - 1
    owner = ctx.owner(O, (j, j))
    acc = torch.zeros(wl.b, wl.b, dtype=torch.float64, device=ctx.device)
    for i in range(j + 1):
        src = ctx.owner(O, (j, i))
        t = O._get_block_ref(j, i) if src == ctx.rank else torch.empty(wl.b, wl.b, dtype=torch.float64, device=ctx.device)
        ctx.bcast_tile(t, src)
        if ctx.rank == owner:
            wl.kernels._gemm_into(acc, acc, t, t, False, True, 1.0, 1.0)
    val = torch.zeros(1, dtype=torch.float64, device=ctx.device)
    if ctx.rank == owner:
        ref = wl.tile(j, j)
        val[0] = (acc - ref).norm() / ref.norm()
    ctx.bcast_tile(val, owner)
    return float(val.item())


def parity_vs_oracle(ctx, n, b, streams, oracle_L, host_tiles):
    """A whole factorisation at the benchmark tile on THIS process grid against the CPU oracle's factor of the same host
    tiles: ||L_gpu - L_oracle||_F / ||L_oracle||_F over every lower tile.  Rank 0 holds the oracle's tiles; every rank
    rebuilds the (deterministic) input tiles it owns."""
    from numpywren_b200.alg_wrappers import cholesky
    from numpywren_b200.matrix import BigMatrix
    nb = n // b
    A = BigMatrix(f"bench_parity_{os.getpid()}", shape=(n, n), shard_sizes=(b, b), device=ctx.device)
    X = None
    for j in range(nb):
        for k in range(j + 1):
            if not ctx.is_mine(A, (j, k)):
                continue
            if host_tiles is not None:
                t = host_tiles[(j, k)]
            else:
                X = X if X is not None else spd_panels(nb, b)
                t = spd_host_tile(X, j, k, n)
            A.put_block(torch.from_numpy(np.ascontiguousarray(t)), j, k)
    program, meta = cholesky(A)
    run_program(ctx, program, streams)
    O = meta["outputs"][0]
    num = den = 0.0
    worst = 0.0
    for j in range(nb):
        for k in range(j + 1):
            src = ctx.owner(O, (j, k))
            t = O._get_block_ref(j, k) if src == ctx.rank else torch.empty(b, b, dtype=torch.float64, device=ctx.device)
            ctx.bcast_tile(t, src)
            if ctx.rank == 0:
                ref = torch.from_numpy(oracle_L[(j, k)]).to(ctx.device)
                d2, r2 = float(((t - ref) ** 2).sum()), float((ref ** 2).sum())
                num += d2; den += r2
                worst = max(worst, (d2 / r2) ** 0.5 if r2 > 0 else 0.0)
    free_all(A, *meta["outputs"], *meta["intermediates"])
    if ctx.rank != 0:
        return None
    return {"rel_fro": (num / den) ** 0.5, "worst_tile_rel": worst, "n": n, "tile": b, "tiles_compared": nb * (nb + 1) // 2,
            "bar": PARITY_BAR, "ok": bool((num / den) ** 0.5 <= PARITY_BAR),
            "what": "||L_gpu - L_oracle||_F / ||L_oracle||_F, whole algs.CHOLESKY program on this process grid vs oracle.run_cholesky on the same host tiles"}


def cholesky_e2e(ctx, wl, streams, steps):
    """Host tiles -> HBM -> factorise -> host tiles on every rank's share, everything inside the timed region, through
    the public API: BigMatrix.put_block(non_blocking=True) from pinned memory (asynchronous H2D on the upload stream,
    consumers wait per tile), BigMatrix.mirror_to_host (every factor tile is copied to pinned host memory as it is
    produced), wait_mirror.  Returns the e2e dict."""
    from numpywren_b200 import job_runner
    from numpywren_b200 import lambdapack as lp
    from numpywren_b200.alg_wrappers import cholesky
    n, b = wl.n, wl.b
    tile_bytes = b * b * 8
    probe = wl.new_matrix("own")
    mine = [(j, k) for (j, k) in wl.lower() if ctx.is_mine(probe, (j, k))]
    cap = host_pinned_cap_bytes(ctx)
    in_bytes = len(mine) * tile_bytes
    ok, err = 1, None
    host_in, ring = {}, []
    n_out = len(mine)
    try:
        if in_bytes + 8 * tile_bytes > cap:
            raise MemoryError(f"pinned input staging {in_bytes >> 30} GiB exceeds this rank's cap {cap >> 30} GiB")
        n_out = int(min(len(mine), max(8, (cap - in_bytes) // tile_bytes)))
        for (j, k) in mine:
            h = torch.empty(b, b, dtype=torch.float64, pin_memory=True)
            h.copy_(wl.tile(j, k))
            host_in[(j, k)] = h
        ring = [torch.empty(b, b, dtype=torch.float64, pin_memory=True) for _ in range(n_out)]
    except Exception as ex:  # pragma: no cover
        ok, err = 0, repr(ex)
    if ctx.min_int(ok) == 0:
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "error": err or "another rank could not stage its host buffers"}
    # every factor tile is read back; when the pinned budget cannot hold the whole factor next to the input, the D2H
    # destination is a ring of pinned tiles (copies to one slot are ordered on the download stream)
    host_out = {jk: ring[i % n_out] for i, jk in enumerate(sorted(mine, key=lambda jk: (jk[1], jk[0])))}
    torch.cuda.synchronize()

    def one():
        torch.cuda.synchronize()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        A = wl.new_matrix("e2e")
        # column by column = the order of first use; the engine waits per tile, so the upload overlaps the factorisation
        for (j, k) in sorted(host_in, key=lambda jk: (jk[1], jk[0])):
            A.put_block(host_in[(j, k)], j, k, non_blocking=True)
        program, meta = cholesky(A)
        O = meta["outputs"][0]
        O.mirror_to_host(host_out)
        program.start()
        job_runner.lambdapack_run(program, timeout=3600, streams=streams, consume_inputs=True)
        O.wait_mirror()
        e1.record()
        e1.synchronize()
        good = program.program_status() == lp.PS.SUCCESS
        ms = ctx.max_float(e0.elapsed_time(e1) if good else float("inf"))
        free_all(A, *meta["outputs"], *meta["intermediates"])
        return ms

    # one warm run, then up to two timed ones; when the wall-clock budget is short the first run is the timed one
    warm = one()
    est = warm * 1e-3
    reps = int(ctx.min_int(max(0, min(steps, 2, int((budget_left() - 60.0) // max(est, 1e-3))))))
    ts = [one() for _ in range(reps)]
    timed = ts if ts else [warm]
    if not np.isfinite(np.mean(timed)):
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": "e2e did not complete"}
    e_ms = float(np.mean(timed))
    moved = ctx.sum_int(in_bytes)
    return {"value": chol_flops(n) / (e_ms * 1e-3) * 1e-12, "unit": UNIT, "h2d_bytes_per_step": moved,
            "d2h_bytes_per_step": moved, "ms_per_step": e_ms, "timed_steps": len(timed), "warm_steps": 1 if ts else 0,
            "d2h_destination": "whole factor resident in pinned host memory" if n_out == len(mine)
            else f"ring of {n_out} pinned tiles per rank (pinned budget {cap >> 30} GiB per rank; every factor tile is still read back)",
            "what": "per rank: pinned host lower tiles -> BigMatrix.put_block(non_blocking=True) (async H2D) -> cholesky() -> "
                    "lambdapack_run -> factor tiles written through to pinned host (D2H) -> wait_mirror; barrier both sides, max over ranks"}


def utilisation_trace(ctx, wl, streams, slices=20):
    """One profiled step (per-node CUDA events): fraction of each time slice this rank's GPU spent inside tile kernels
    (union of node intervals), from hardware timestamps.  -> per-rank list, gathered on rank 0."""
    from numpywren_b200 import job_runner
    from numpywren_b200.alg_wrappers import cholesky
    A = wl.resident_input()
    program, meta = cholesky(A)
    ms, _, _ = run_program(ctx, program, streams, profile=True)
    tl = job_runner.node_timeline(program)
    iv = sorted((s, e) for _, _, s, e, _ in tl)
    end = max((e for _, e in iv), default=0.0)
    span = ctx.max_float(end)
    busy = [0.0] * slices
    cur_s, cur_e = None, None
    merged = []
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                merged.append((cur_s, cur_e))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        merged.append((cur_s, cur_e))
    w = max(span / slices, 1e-9)
    for s, e in merged:
        for i in range(slices):
            lo, hi = i * w, (i + 1) * w
            busy[i] += max(0.0, min(e, hi) - max(s, lo))
    util = [round(x / w, 3) for x in busy]
    free_all(A, *meta["outputs"], *meta["intermediates"])
    out = [None] * ctx.world
    if ctx.world > 1:
        import torch.distributed as dist
        dist.all_gather_object(out, util)
    else:
        out = [util]
    by_kernel = {}
    for name, _, s, e, _ in tl:
        by_kernel[name] = by_kernel.get(name, 0.0) + (e - s)
    return {"step_ms": ms, "slices": slices, "busy_fraction_per_rank": out,
            "mean_over_ranks": [round(float(np.mean([r[i] for r in out])), 3) for i in range(slices)],
            "rank0_kernel_ms": {k: round(v, 1) for k, v in by_kernel.items()}}


def run_cholesky_arm(args):
    ctx = Ctx()
    n, b = args.n, args.tile
    wl = CholeskyWorkload(ctx, n, b)
    streams = args.streams
    c, r, s = chol_task_counts(wl.nb)

    # ---- second labelled workload at N=1: BASELINE configs[1] (N=65536 on a single B200)
    also = None
    if ctx.world == 1 and n != 65536 and args.tile == 4096 and not args.no_also and budget_left() > 600:
        wl2 = CholeskyWorkload(ctx, 65536, b)
        t2 = []
        for i in range(3):
            ms2, _, _, A, program, meta = cholesky_step(ctx, wl2, streams)
            free_all(A, *meta["outputs"], *meta["intermediates"])
            del A, program, meta
            if i > 0:
                t2.append(ms2)
        also = {"workload": "Cholesky N=65536 fp64 SPD, tile=4096, single B200 (BASELINE configs[1])", "ms_per_step": float(np.mean(t2)),
                "value": chol_flops(65536) / (float(np.mean(t2)) * 1e-3) * 1e-12, "unit": UNIT, "steps": 2, "warmup": 1}
        del wl2

    # ---- warm-up
    resid, expand_s, plan_s = None, None, None
    for w in range(args.warmup):
        ms, _, plan_s, A, program, meta = cholesky_step(ctx, wl, streams)
        if ctx.rank == 0:
            log(f"warmup {w}: {ms:.1f} ms")
        if w == args.warmup - 1:
            resid = residual_check(ctx, wl, meta["outputs"][0])
            expand_s = program.program.expand_time
        free_all(A, *meta["outputs"], *meta["intermediates"])
        del A, program, meta
    # ---- timed steps
    sampler = ClockSampler(index=ctx.local_rank).start() if ctx.rank == 0 else None
    times, launches_tot, sent = [], 0, 0
    for _ in range(args.steps):
        ms, launches, plan_s, A, program, meta = cholesky_step(ctx, wl, streams)
        times.append(ms)
        launches_tot += launches
        eng = program._engine
        sent = eng.comm.bytes_sent if eng.comm is not None else 0
        if resid is None:
            resid = residual_check(ctx, wl, meta["outputs"][0])
            expand_s = program.program.expand_time
        free_all(A, *meta["outputs"], *meta["intermediates"])
        del A, program, meta
    clocks = sampler.stop() if sampler is not None else None
    ms_per_step = float(np.mean(times))
    value = chol_flops(n) / (ms_per_step * 1e-3) * 1e-12
    nvlink = ctx.sum_int(sent)
    if ctx.rank == 0:
        log(f"timed: {ms_per_step:.1f} ms/step -> {value:.2f} TFLOP/s; budget left {budget_left():.0f} s")

    # ---- dominant kernel alone (roofline), rank 0
    roofline = tensor_roofline(ctx, b, value) if ctx.rank == 0 else None

    # ---- optional per-slice utilisation trace from hardware timestamps (one extra profiled step)
    trace = None
    if args.trace:
        trace = utilisation_trace(ctx, wl, streams)

    # ---- end to end with host buffers at this GPU count
    if os.environ.get("NPW_B200_BENCH_NO_E2E"):
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": "skipped by NPW_B200_BENCH_NO_E2E"}
    else:
        try:
            e2e = cholesky_e2e(ctx, wl, streams, args.steps)
        except Exception as ex:  # pragma: no cover
            log("e2e failed:", repr(ex))
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)}
    del wl.X

    # ---- CPU baseline (rank 0, bounded sample) + parity of a whole factorisation against the oracle's factor
    cpu, parity, oracle_L, host_tiles = None, None, None, None
    do_cpu = not args.no_cpu
    if do_cpu and ctx.rank == 0:
        ks = CpuKernelSample(b)
        ks.once()
        reps = [ks.once() for _ in range(3)]
        mean = {k: float(np.mean([x[k] for x in reps])) for k in ("syrk", "trsm", "chol")}
        t_comp = CpuKernelSample.compose(mean, wl.nb)
        t_meas, oracle_L, host_tiles, cores = cpu_cholesky_sample(args.cpu_n, b)
        cpu = {"value": chol_flops(n) / t_comp * 1e-12, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": (f"oracle syrk/trsm/chol at {b}x{b} on {cores} host threads (3 samples each: "
                          f"{mean['syrk']:.3f}/{mean['trsm']:.3f}/{mean['chol']:.3f} s), composed by the task counts of N={n} "
                          f"({s} syrk + {r} trsm + {c} chol = {t_comp:.0f} s, EXTRAPOLATED); cross-check: oracle run_cholesky "
                          f"N={args.cpu_n} measured {t_meas:.1f} s vs {CpuKernelSample.compose(mean, args.cpu_n // b):.1f} s composed"),
               "kernel_seconds": mean,
               "measured_sample": {"n": args.cpu_n, "seconds": t_meas, "tflops": chol_flops(args.cpu_n) / t_meas * 1e-12},
               "config1_gemm": cpu_config1_gemm()}
    if do_cpu:
        parity = parity_vs_oracle(ctx, args.cpu_n, b, streams, oracle_L, host_tiles)

    if ctx.rank == 0:
        grid = ctx.grid
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": _dtype_label(), "data": "synthetic",
                "config": {"workload": workload_name(args), "tile_tasks": c + r + s,
                           "process_grid": (f"{grid.P}x{grid.Q} rotated block-cyclic over tile index" if grid else "1x1"),
                           "exchange": os.environ.get("NPW_B200_EXCHANGE", "symm") if grid else None,
                           "l2": "inputs larger than L2 (every step regenerates its %.1f GB of input tiles)" % (
                               wl.nb * (wl.nb + 1) // 2 * b * b * 8 / 1e9),
                           "streams": streams, "dag_expand_s": expand_s, "plan_s": plan_s,
                           "nvlink_bytes_per_step": nvlink,
                           "parity_vs_oracle": parity, "residual_LLt_minus_A": resid,
                           "algorithmic_flops_per_step": chol_flops(n), "also": also, "utilisation_trace": trace},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_tot), "clocks": clocks}
        print(json.dumps(line), flush=True)
    ctx.finish()
    if parity is not None and not parity["ok"]:
        log("PARITY FAILED:", parity)
        return 3
    return 0


# ----------------------------------------------------------------------------------------------- TSQR workload (config 4)
def golden_tsqr_parity(ctx, streams, name="tsqr_256_32"):
    """The TSQR program on THIS process grid against the fixture written by the unmodified reference
    (tests/golden/<name>.npz, oracle/make_golden.py): final R elementwise (same Householder convention, so signs agree)."""
    from numpywren_b200.alg_wrappers import _place_by_row_block, tsqr
    from numpywren_b200.matrix import BigMatrix
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.exists(path):
        return None
    g = np.load(path)
    m, b, nlev = int(g["m"]), int(g["b"]), int(g["nlev"])
    X = BigMatrix(f"bench_tsqr_golden_{os.getpid()}", shape=(m, b), shard_sizes=(b, b), device=ctx.device)
    X.free()
    _place_by_row_block(X, 0)
    for j in range(m // b):
        if ctx.is_mine(X, (j, 0)):
            X.put_block(torch.from_numpy(np.ascontiguousarray(g["X"][j * b:(j + 1) * b])), j, 0)
    program, meta = tsqr(X)
    run_program(ctx, program, streams, consume=False)
    Rs = meta["outputs"][0]
    src = ctx.owner(Rs, (nlev, 0))
    R = Rs._get_block_ref(nlev, 0).reshape(b, b).contiguous() if src == ctx.rank else torch.empty(b, b, dtype=torch.float64, device=ctx.device)
    R = ctx.bcast_tile(R, src)
    ref = torch.from_numpy(g["R"]).to(ctx.device)
    err = float((R - ref).norm() / ref.norm())
    free_all(X, *meta["outputs"])
    return {"fixture": f"tests/golden/{name}.npz (written by the unmodified reference)", "rel_fro_R": err, "bar": PARITY_BAR,
            "ok": bool(err <= PARITY_BAR), "leaves": m // b, "what": "final R of algs.TSQR on this process grid vs the reference's"}


def golden_gemm_parity(ctx, streams, name="gemm_64_16"):
    """The GEMM program on THIS process grid against the fixture written by the unmodified reference."""
    from numpywren_b200 import alg_wrappers
    from numpywren_b200.matrix import BigMatrix
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.exists(path):
        return None
    g = np.load(path)
    n, b = int(g["n"]), int(g["b"])
    nb = n // b
    mats = []
    for tag in ("A", "B"):
        M = BigMatrix(f"bench_gemm_golden_{tag}_{os.getpid()}", shape=(n, n), shard_sizes=(b, b), device=ctx.device)
        M.free()
        for i in range(nb):
            for j in range(nb):
                if ctx.is_mine(M, (i, j)):
                    M.put_block(torch.from_numpy(np.ascontiguousarray(g[tag][i * b:(i + 1) * b, j * b:(j + 1) * b])), i, j)
        mats.append(M)
    program, meta = alg_wrappers.gemm_kloop(mats[0], mats[1], out_key=f"bench_gemm_golden_C_{os.getpid()}")
    run_program(ctx, program, streams, consume=False)
    C = meta["outputs"][0]
    num = den = 0.0
    for i in range(nb):
        for j in range(nb):
            src = ctx.owner(C, (i, j))
            t = C._get_block_ref(i, j).reshape(b, b).contiguous() if src == ctx.rank else torch.empty(b, b, dtype=torch.float64, device=ctx.device)
            t = ctx.bcast_tile(t, src)
            ref = torch.from_numpy(np.ascontiguousarray(g["C"][i * b:(i + 1) * b, j * b:(j + 1) * b])).to(ctx.device)
            num += float(((t - ref) ** 2).sum()); den += float((ref ** 2).sum())
    free_all(*mats, *meta["outputs"], *meta.get("intermediates", []))
    err = (num / den) ** 0.5
    return {"fixture": f"tests/golden/{name}.npz (written by the unmodified reference)", "rel_fro_C": err, "bar": PARITY_BAR,
            "ok": bool(err <= PARITY_BAR), "what": "C of the GEMM_ACC program on this process grid vs the reference's algs.GEMM result"}


def run_tsqr_arm(args):
    """BASELINE config 4: TSQR m x 512 fp64, tile (65536, 512), binary reduction tree; row blocks sharded over the GPUs
    (block j on rank j mod world), tree merges exchange 2 MiB R factors over NVLink."""
    from numpywren_b200 import kernels
    from numpywren_b200.alg_wrappers import _place_by_row_block, tsqr
    from numpywren_b200.matrix import BigMatrix
    ctx = Ctx()
    m, ncol, b = args.n, args.cols, args.tile
    nb = m // b
    levels = max(int(np.ceil(np.log2(nb))), 1)
    flops = 2.0 * m * ncol * ncol - 2.0 * ncol ** 3 / 3.0
    step_id = [0]

    def make_input():
        step_id[0] += 1
        X = BigMatrix(f"bench_tsqr_{step_id[0]}", shape=(m, ncol), shard_sizes=(b, ncol), device=ctx.device)
        X.free()
        _place_by_row_block(X, 0)
        for j in range(nb):
            if ctx.is_mine(X, (j, 0)):
                t = torch.empty(b, ncol, dtype=torch.float64, device=ctx.device)
                kernels.fill_random(t, seed=3, row0=j * b)
                X._put_block_ref(t, j, 0)
        return X

    def gram_check(X, R):
        """||R^T R - X^T X||_F / ||X^T X||_F: the size-independent property (R is unique up to row signs)."""
        G = torch.zeros(ncol, ncol, dtype=torch.float64, device=ctx.device)
        for j in range(nb):
            if ctx.is_mine(X, (j, 0)):
                t = X._get_block_ref(j, 0)
                G += t.T @ t                      # checker arithmetic (cuBLAS), outside the timed region
        if ctx.world > 1:
            import torch.distributed as dist
            dist.all_reduce(G)
        return float((R.T @ R - G).norm() / G.norm())

    timeline = [None]

    def one(check=False, profile=False):
        X = make_input()
        program, meta = tsqr(X)
        ms, launches, plan_s = run_program(ctx, program, args.streams, consume=False, profile=profile)
        if profile:
            from numpywren_b200 import job_runner
            tl = [[name, int(v.get("j", -1)), int(v.get("level", -1)), round(s0, 2), round(e0, 2)]
                  for name, v, s0, e0, _ in job_runner.node_timeline(program)]
            out = [None] * ctx.world
            if ctx.world > 1:
                import torch.distributed as dist
                dist.all_gather_object(out, tl)
            else:
                out = [tl]
            timeline[0] = {"step_ms": ms, "per_rank_nodes": out, "columns": ["kernel", "j", "level", "start_ms", "end_ms"]}
        err = None
        if check:
            Rs = meta["outputs"][0]
            src = ctx.owner(Rs, (levels, 0))
            R = Rs._get_block_ref(levels, 0) if src == ctx.rank else torch.empty(ncol, ncol, dtype=torch.float64, device=ctx.device)
            R = ctx.bcast_tile(R.reshape(ncol, ncol).contiguous(), src)
            err = gram_check(X, R)
        nodes = len(program.program.nodes)
        free_all(X, *meta["outputs"], *meta.get("intermediates", []))
        return ms, launches, err, nodes

    err, nodes = None, 0
    for w in range(args.warmup):
        ms, _, e, nodes = one(check=(w == args.warmup - 1))
        err = e if e is not None else err
        if ctx.rank == 0:
            log(f"warmup {w}: {ms:.1f} ms")
    sampler = ClockSampler(index=ctx.local_rank).start() if ctx.rank == 0 else None
    times, launches_tot = [], 0
    for _ in range(args.steps):
        ms, launches, _, nodes = one()
        times.append(ms); launches_tot += launches
    clocks = sampler.stop() if sampler is not None else None
    ms_per_step = float(np.mean(times))
    value = flops / (ms_per_step * 1e-3) * 1e-12
    if args.trace:
        one(profile=True)
    parity = golden_tsqr_parity(ctx, args.streams)
    roofline = None
    if ctx.rank == 0:
        peaks = measure_peaks(ctx.device)
        peak = max(peaks.values())
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6549.1
        # the leaf factorisation alone on rank 0 (one 65536 x 512 row block)
        a = torch.empty(b, ncol, dtype=torch.float64, device=ctx.device); kernels.fill_random(a, 5)
        kernels.qr_factor(a); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); kernels.qr_factor(a); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        leaf_ms = float(np.mean(ts))
        leaf_flops = 2.0 * b * ncol * ncol - 2.0 * ncol ** 3 / 3.0
        ach = leaf_flops / (leaf_ms * 1e-3) * 1e-12
        gbs = 2.0 * b * ncol * 8 / (leaf_ms * 1e-3) * 1e-9
        roofline = {"bound": "tensor", "kernel": "npw_geqrt_f64 (kernels.qr_factor on one %dx%d leaf: panel + compact-WY kernels)" % (b, ncol),
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                    "flops_per_launch": leaf_flops, "avg_launch_ms": leaf_ms,
                    "hbm": {"achieved_gbs": gbs, "peak_gbs": hbm, "frac": gbs / hbm, "bytes": 2.0 * b * ncol * 8,
                            "what": "read A + write V of the leaf (algorithmic bytes) / leaf time"},
                    "whole_job_frac_of_aggregate_peak": value / (peak * ctx.world),
                    "whole_job_hbm_frac": (2.0 * m * ncol * 8 / (ms_per_step * 1e-3) * 1e-9) / (hbm * ctx.world),
                    "peak_source": "fp64 pipe measured live on rank 0 (%.2f TFLOP/s); HBM from MEASURED_PEAKS.json" % peak}
    if ctx.rank == 0:
        line = {"metric": f"fp64 TFLOP/s TSQR {m}x{ncol} tile=({b},{ncol})", "value": value, "unit": UNIT, "n_gpus": ctx.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args), "tile_tasks": nodes, "tree_levels": levels,
                           "placement": "leaf j on rank j mod world; merge k of every tree level on rank k mod world", "streams": args.streams,
                           "gram_residual_RtR_minus_XtX": err, "parity_vs_golden": parity, "algorithmic_flops_per_step": flops,
                           "node_timeline": timeline[0],
                           "parity": "kernels vs LAPACK restatement at 1e-10 in tests/ (QR family: parity unpinned beyond LAPACK, DESIGN §7)",
                           "l2": "inputs larger than L2 (%.1f GB per step)" % (m * ncol * 8 / 1e9)},
                "roofline": roofline, "cpu_baseline": None,
                "e2e": {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "host-buffer end-to-end is reported for the headline Cholesky workload"},
                "gpu_launches": int(launches_tot), "clocks": clocks}
        print(json.dumps(line), flush=True)
    ctx.finish()
    return 0


# ----------------------------------------------------------------------------------------------- GEMM workload (config 5)
def run_gemm_arm(args):
    """BASELINE config 5: C = A B, N x N fp64, tile 8192, the legacy binops.gemm owner-computes / serial-K schedule as
    the LambdaPACK program GEMM_ACC (A tiles travel along process rows, B tiles along process columns)."""
    from numpywren_b200 import alg_wrappers, kernels
    from numpywren_b200.matrix import BigMatrix
    ctx = Ctx()
    n, b = args.n, args.tile
    nb = n // b
    flops = 2.0 * n ** 3
    A = BigMatrix("bench_gemm_A", shape=(n, n), shard_sizes=(b, b), device=ctx.device); A.free()
    B = BigMatrix("bench_gemm_B", shape=(n, n), shard_sizes=(b, b), device=ctx.device); B.free()
    alg_wrappers.place_plain_block_cyclic(A)       # before the tiles are created: ownership decides where they live
    alg_wrappers.place_plain_block_cyclic(B)
    for i in range(nb):
        for k in range(nb):
            for mtx, seed in ((A, 1), (B, 2)):
                if ctx.is_mine(mtx, (i, k)):
                    t = torch.empty(b, b, dtype=torch.float64, device=ctx.device)
                    kernels.fill_random(t, seed, i * b, k * b)
                    mtx._put_block_ref(t, i, k)
    rep = [0]

    def one(check=False):
        rep[0] += 1
        program, meta = alg_wrappers.gemm_kloop(A, B, out_key=f"bench_gemm_C{rep[0]}")
        ms, launches, plan_s = run_program(ctx, program, args.streams, consume=False)
        err = None
        if check:
            # C[0, nb-1] against a plain K-loop over gathered tiles with torch.matmul (cuBLAS: an independent arithmetic
            # check of schedule, exchange and kernel), outside the timed region
            C = meta["outputs"][0]
            i, j = 0, nb - 1
            acc = torch.zeros(b, b, dtype=torch.float64, device=ctx.device)
            owner = ctx.owner(C, (i, j))
            for k in range(nb):
                ta = A._get_block_ref(i, k) if ctx.is_mine(A, (i, k)) else torch.empty(b, b, dtype=torch.float64, device=ctx.device)
                tb = B._get_block_ref(k, j) if ctx.is_mine(B, (k, j)) else torch.empty(b, b, dtype=torch.float64, device=ctx.device)
                ctx.bcast_tile(ta, ctx.owner(A, (i, k))); ctx.bcast_tile(tb, ctx.owner(B, (k, j)))
                if ctx.rank == owner:
                    acc += ta @ tb
            v = torch.zeros(1, dtype=torch.float64, device=ctx.device)
            if ctx.rank == owner:
                v[0] = (C._get_block_ref(i, j) - acc).norm() / acc.norm()
            ctx.bcast_tile(v, owner)
            err = float(v.item())
        nodes = len(program.program.nodes)
        free_all(*meta["outputs"], *meta["intermediates"])
        return ms, launches, err, nodes

    err, nodes = None, 0
    for w in range(args.warmup):
        ms, _, e, nodes = one(check=(w == args.warmup - 1))
        err = e if e is not None else err
        if ctx.rank == 0:
            log(f"warmup {w}: {ms:.1f} ms")
    sampler = ClockSampler(index=ctx.local_rank).start() if ctx.rank == 0 else None
    times, launches_tot = [], 0
    for _ in range(args.steps):
        ms, launches, _, nodes = one()
        times.append(ms); launches_tot += launches
    clocks = sampler.stop() if sampler is not None else None
    ms_per_step = float(np.mean(times))
    value = flops / (ms_per_step * 1e-3) * 1e-12
    free_all(A, B)
    parity = golden_gemm_parity(ctx, args.streams)
    roofline = tensor_roofline(ctx, b, value) if ctx.rank == 0 else None
    if ctx.rank == 0:
        grid = ctx.grid
        line = {"metric": f"fp64 TFLOP/s GEMM N={n} tile={b}", "value": value, "unit": UNIT, "n_gpus": ctx.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args), "tile_tasks": nodes, "streams": args.streams,
                           "process_grid": (f"{grid.P}x{grid.Q} plain block-cyclic over tile index" if grid else "1x1"),
                           "tile_rel_err_vs_cublas_kloop": err, "parity_vs_golden": parity, "algorithmic_flops_per_step": flops,
                           "l2": "inputs larger than L2 (%.1f GB of A and B tiles)" % (2 * n * n * 8 / 1e9)},
                "roofline": roofline, "cpu_baseline": None,
                "e2e": {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "host-buffer end-to-end is reported for the headline Cholesky workload"},
                "gpu_launches": int(launches_tot), "clocks": clocks}
        print(json.dumps(line), flush=True)
    ctx.finish()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cholesky", choices=["cholesky", "tsqr", "gemm"])
    ap.add_argument("--size", dest="n", type=int, default=None,
                    help="cholesky/gemm: matrix size N (default 131072; gemm on fewer than 8 GPUs: 65536); tsqr: rows (default 4194304)")
    ap.add_argument("--tile", type=int, default=None, help="tile size (cholesky 4096, gemm 8192, tsqr 65536 rows)")
    ap.add_argument("--cols", type=int, default=512, help="tsqr: columns")
    ap.add_argument("--streams", type=int, default=None,
                    help="compute streams per GPU (default 4 on one GPU, 8 on several: fewer head-of-line stalls on remote tiles)")
    ap.add_argument("--cpu-n", dest="cpu_n", type=int, default=16384, help="size of the measured CPU sample / parity run")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline and the oracle parity run")
    ap.add_argument("--no-also", action="store_true", help="skip the second labelled N=65536 measurement at one GPU")
    ap.add_argument("--trace", action="store_true", help="add a per-slice GPU utilisation trace (one extra profiled step)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.tile is None:
        args.tile = {"cholesky": 4096, "gemm": 8192, "tsqr": 65536}[args.workload]
    if args.n is None:
        args.n = {"cholesky": 131072, "gemm": 131072 if world >= 8 else 65536, "tsqr": 4194304}[args.workload]
    if args.streams is None:
        args.streams = 4 if world == 1 else 8
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "tsqr":
        return run_tsqr_arm(args)
    if args.workload == "gemm":
        return run_gemm_arm(args)
    return run_cholesky_arm(args)


if __name__ == "__main__":
    sys.exit(main())
