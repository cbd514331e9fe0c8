"""Discrete-event model of the multi-GPU Cholesky schedule (analysis tool, not product code).

What is REAL here: the expanded DAG (compiler), the host's enqueue order (the actual LambdaPackProgram ready queue with
the engine's critical-path priorities, driven by job_runner.lambdapack_run with a recording engine), the stream chosen
for every node (TileEngine._stream_for's hash), the tile ownership map and the transfer plan (parallel.TransferPlan).
What is MODELLED: the device.  Per GPU, every stream is a FIFO; a kernel may start when it is at the head of its
stream and all its input tiles have arrived; ready kernels share the GPU by priority-then-FIFO processor sharing: a
kernel has a latency L and an amount of whole-GPU work W <= L (a tile GEMM has W = L; the tile Cholesky is latency
bound, W << L), it progresses at rate <= W/L of the GPU, high-priority streams are served first, normal streams in
the order their head became ready.  A produced tile that another rank needs is copied on a per-(src,dst) FIFO link
at `--link-gbs`; all links of one source share `--egress-gbs`.

  python tools/schedule_sim.py --nb 32 --grid 2x4 --streams 8
prints the simulated makespan, the fp64 efficiency against the ideal (sum of W / ranks), and where the idle time sits.
Durations default to the kernel timings measured on B200 in round 1; they are inputs, not results.
"""
from __future__ import annotations

import argparse
import heapq
import os
import sys
from collections import defaultdict, deque

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from numpywren_b200 import alg_wrappers, job_runner, parallel  # noqa: E402
from numpywren_b200.compiler import _tile_key  # noqa: E402
from numpywren_b200.matrix import BigMatrix  # noqa: E402

TILE_BYTES = 4096 * 4096 * 8


def build(nb, P, Q, owner_fn=None):
    """-> (compiled, enqueue order [nid], exec_rank[nid], transfers after node {nid: [(key, src, dst)]}, prio[nid])."""
    A = BigMatrix("sim_A_%d" % nb, shape=(nb * 4, nb * 4), shard_sizes=(4, 4), device="cpu")
    program, meta = alg_wrappers.cholesky(A)
    compiled = program.program
    order = []

    class Recorder(job_runner.TileEngine):
        def run_node(self, node):
            order.append(node.nid)

        def finish(self):
            return []

    eng = Recorder(program)
    program._engine = eng
    prio = eng.priorities()
    program._priority_fn = lambda e, v: prio[compiled.node(e, v).nid]
    program.start()
    job_runner.lambdapack_run(program, timeout=3600)
    grid = parallel.ProcessGrid(P * Q, 0, shape=(P, Q))
    if owner_fn is not None:
        grid.owner = lambda m, idx: owner_fn(grid, *grid.coords(m, idx))
    plan = parallel.TransferPlan(compiled, grid)
    after = {nid: [(k, s, d) for (k, _, _, s, d) in lst] for nid, lst in plan.after_node.items()}
    # the send stream of a (src, dst) pair is a FIFO in HOST order: position of every transfer on its link
    link_seq = {}
    counters = defaultdict(int)
    for nid in order:
        for (k, s_, d) in after.get(nid, ()):
            link_seq[(k, d)] = counters[(s_, d)]
            counters[(s_, d)] += 1
    return compiled, order, plan.exec_rank, after, prio, grid, link_seq


def stream_of(node, n_streams, n_high):
    """TileEngine._stream_for: hash of the output tile's last two block coordinates; chol/trsm go to the high pool."""
    m, idx = node.writes[0]
    tail = idx[-2:] if len(idx) == 3 else idx
    h = 0
    for v in tail:
        h = h * 1000003 + int(v) + 7
    if node.call.compute_name in job_runner.HIGH_PRIORITY_KERNELS and n_high:
        return ("H", h % n_high)
    return ("N", h % n_streams)


def simulate(nb, P, Q, n_streams, n_high, t, link_gbs, egress_gbs, owner_fn=None, policy="fifo", verbose=False,
             critical_high=False):
    compiled, order, exec_rank, after, prio, grid, link_seq = build(nb, P, Q, owner_fn)
    nodes = compiled.nodes
    world = P * Q
    # ---- per-node duration model
    dur = {}
    for n in nodes:
        name = n.call.compute_name
        if name == "chol":
            dur[n.nid] = (t["potrf_L"], t["potrf_W"])
        elif name == "trsm":
            dur[n.nid] = (t["trsm_L"], t["trsm_W"])
        else:
            diag = _tile_key(*n.reads[1]) == _tile_key(*n.reads[2])
            L = t["syrk_diag"] if diag else t["syrk"]
            dur[n.nid] = (L, L)
    # ---- streams: FIFO of nids per (rank, stream) in host enqueue order
    fifo = defaultdict(deque)
    stream_id = {}
    for nid in order:
        n = nodes[nid]
        s = stream_of(n, n_streams, n_high)
        if critical_high and n.call.compute_name == "syrk" and n_high:
            # the update that feeds the next panel's chol/trsm directly (tile column == version) is critical too
            m, idx = n.writes[0]
            if idx[0] == idx[2]:
                s = ("H", s[1] % n_high)
        stream_id[nid] = (exec_rank[nid], s)
        fifo[(exec_rank[nid], s)].append(nid)
    # ---- dependencies: tile -> time available on a rank
    writer = {}
    for n in nodes:
        for (m, idx) in n.writes:
            writer[_tile_key(m, idx)] = n.nid
    avail = {}                     # (tile_key, rank) -> time
    need = {}                      # nid -> list of (tile_key) produced by other nodes
    for n in nodes:
        need[n.nid] = [_tile_key(m, idx) for (m, idx) in n.reads if _tile_key(m, idx) in writer]
    # ---- event loop
    now = 0.0
    finish_time = {}
    running = {r: {} for r in range(world)}          # rank -> {nid: [remaining_work, rate_cap, ready_time, high]}
    busy = [0.0] * world
    link_free = defaultdict(float)                   # (src, dst) -> time
    egress_free = [0.0] * world
    arrivals = []                                    # heap of (time, key, dst)
    pending = defaultdict(dict)                      # (src, dst) -> {position on the link: (key, producer finish time)}
    link_next = defaultdict(int)
    slices = 20
    util_t = []                                      # (t0, t1, total service rate over all ranks)
    heads_checked = True

    def deps_ready(nid, r):
        tmax = 0.0
        for k in need[nid]:
            tt = avail.get((k, r))
            if tt is None:
                return None
            tmax = max(tmax, tt)
        return tmax

    def try_start(r):
        started = False
        for (rr, s), q in fifo.items():
            if rr != r or not q:
                continue
            nid = q[0]
            if nid in running[r]:
                continue
            # a stream runs one kernel at a time
            if any(stream_id[x] == (r, s) for x in running[r]):
                continue
            tr = deps_ready(nid, r)
            if tr is None or tr > now + 1e-12:
                continue
            L, W = dur[nid]
            running[r][nid] = [W, W / L, now, s[0] == "H"]
            started = True
        return started

    def rates(r):
        """Processor sharing: high-priority kernels first (each capped at W/L), then normal ones FIFO by ready time
        (policy 'fifo') or equally (policy 'share')."""
        cap = 1.0
        out = {}
        items = sorted(running[r].items(), key=lambda kv: (not kv[1][3], kv[1][2], kv[0]))
        if policy == "share":
            hi = [kv for kv in items if kv[1][3]]
            lo = [kv for kv in items if not kv[1][3]]
            for nid, st in hi:
                g = min(st[1], cap)
                out[nid] = g
                cap -= g
            if lo:
                # equal split, respecting caps
                rem = list(lo)
                while rem and cap > 1e-12:
                    share = cap / len(rem)
                    nxt = []
                    used = 0.0
                    for nid, st in rem:
                        g = min(st[1] - out.get(nid, 0.0), share)
                        out[nid] = out.get(nid, 0.0) + g
                        used += g
                        if out[nid] < st[1] - 1e-12:
                            nxt.append((nid, st))
                    cap -= used
                    if used < 1e-12:
                        break
                    rem = nxt
            return out
        for nid, st in items:
            g = min(st[1], cap)
            out[nid] = g
            cap -= g
        return out

    total = len(nodes)
    done = 0
    for r in range(world):
        try_start(r)
    guard = 0
    while done < total:
        guard += 1
        if guard > 50 * total + 1000:
            raise RuntimeError("simulation does not make progress")
        # next event: a kernel finishing or a tile arriving
        t_next = None
        rate_cache = {}
        for r in range(world):
            if not running[r]:
                continue
            rt = rates(r)
            rate_cache[r] = rt
            for nid, st in running[r].items():
                g = rt.get(nid, 0.0)
                if g > 1e-15:
                    tf = now + st[0] / g
                    if t_next is None or tf < t_next:
                        t_next = tf
        if arrivals and (t_next is None or arrivals[0][0] < t_next):
            t_next = arrivals[0][0]
        if t_next is None:
            stuck = [(r, s, q[0], nodes[q[0]]) for (r, s), q in fifo.items() if q]
            raise RuntimeError(f"deadlock at t={now}: {stuck[:4]}")
        dt = t_next - now
        util_t.append((now, t_next, sum(sum(rt.values()) for rt in rate_cache.values())))
        for r in range(world):
            rt = rate_cache.get(r)
            if not rt:
                continue
            busy[r] += dt * sum(rt.values())
            for nid, st in running[r].items():
                st[0] -= dt * rt.get(nid, 0.0)
        now = t_next
        touched = set()
        while arrivals and arrivals[0][0] <= now + 1e-12:
            _, k, d = heapq.heappop(arrivals)
            avail[(k, d)] = now
            touched.add(d)
        for r in range(world):
            fin = [nid for nid, st in running[r].items() if st[0] <= 1e-9]
            for nid in fin:
                del running[r][nid]
                fifo[stream_id[nid]].popleft()
                finish_time[nid] = now
                done += 1
                touched.add(r)
                for (m, idx) in nodes[nid].writes:
                    avail[(_tile_key(m, idx), r)] = now
                for (k, src, dst) in after.get(nid, ()):
                    pending[(src, dst)][link_seq[(k, dst)]] = (k, now)
                    # issue every transfer of this link that is next in host order and whose producer has finished
                    while link_next[(src, dst)] in pending[(src, dst)]:
                        kk, tready = pending[(src, dst)].pop(link_next[(src, dst)])
                        link_next[(src, dst)] += 1
                        start = max(tready, link_free[(src, dst)], egress_free[src] if egress_gbs else 0.0)
                        end = start + TILE_BYTES / (link_gbs * 1e9) * 1e3 + t["signal"]
                        link_free[(src, dst)] = end
                        if egress_gbs:
                            egress_free[src] = start + TILE_BYTES / (egress_gbs * 1e9) * 1e3
                        heapq.heappush(arrivals, (end, kk, dst))
        for r in touched:
            while try_start(r):
                pass
    work = [0.0] * world
    for n in nodes:
        work[exec_rank[n.nid]] += dur[n.nid][1]
    ideal = sum(work) / world
    # fp64 efficiency the way bench.py reports it: N^3/3 flops at the pipe peak vs makespan
    flops = (nb * 4096.0) ** 3 / 3.0
    eff_peak = flops / (now * 1e-3) / (world * t["peak_tflops"] * 1e12)
    res = {"makespan_ms": now, "ideal_ms": ideal, "max_rank_work_ms": max(work), "sched_eff": ideal / now,
           "frac_of_peak": eff_peak, "tflops": flops / (now * 1e-3) / 1e12, "busy_frac": [b / now for b in busy]}
    prof = [0.0] * slices
    for (a, b, u) in util_t:
        for sidx in range(slices):
            lo, hi = now * sidx / slices, now * (sidx + 1) / slices
            ov = max(0.0, min(b, hi) - max(a, lo))
            prof[sidx] += ov * u
    res["util_profile"] = [p / (now / slices) / world for p in prof]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nb", type=int, default=32)
    ap.add_argument("--grid", default="2x4")
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--high", type=int, default=2)
    ap.add_argument("--syrk", type=float, default=3.83)
    ap.add_argument("--syrk-diag", type=float, default=2.10)
    ap.add_argument("--trsm-L", type=float, default=2.9)
    ap.add_argument("--trsm-W", type=float, default=2.3)
    ap.add_argument("--potrf-L", type=float, default=5.7)
    ap.add_argument("--potrf-W", type=float, default=1.5)
    ap.add_argument("--link-gbs", type=float, default=600.0)
    ap.add_argument("--egress-gbs", type=float, default=0.0)
    ap.add_argument("--signal", type=float, default=0.02)
    ap.add_argument("--peak", type=float, default=36.84)
    ap.add_argument("--policy", default="fifo", choices=["fifo", "share"])
    ap.add_argument("--critical-high", action="store_true")
    a = ap.parse_args()
    P, Q = (int(x) for x in a.grid.split("x"))
    t = {"syrk": a.syrk, "syrk_diag": a.syrk_diag, "trsm_L": a.trsm_L, "trsm_W": a.trsm_W, "potrf_L": a.potrf_L,
         "potrf_W": a.potrf_W, "signal": a.signal, "peak_tflops": a.peak}
    r = simulate(a.nb, P, Q, a.streams, a.high, t, a.link_gbs, a.egress_gbs, policy=a.policy, critical_high=a.critical_high)
    print(f"nb={a.nb} grid={P}x{Q} streams={a.streams}+{a.high} policy={a.policy}: makespan {r['makespan_ms']:.0f} ms, "
          f"{r['tflops']:.1f} TFLOP/s = {100 * r['frac_of_peak']:.1f} % of {P * Q} x {a.peak} peak; "
          f"ideal(sum W / ranks) {r['ideal_ms']:.0f} ms, max rank work {r['max_rank_work_ms']:.0f} ms, "
          f"schedule efficiency {100 * r['sched_eff']:.1f} %")
    print("   busy fraction per rank:", " ".join(f"{b:.3f}" for b in r["busy_frac"]))
    print("   utilisation over time (20 slices):", " ".join(f"{u:.2f}" for u in r["util_profile"]))


if __name__ == "__main__":
    main()
