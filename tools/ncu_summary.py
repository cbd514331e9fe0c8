"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_summary.txt
  python tools/ncu_summary.py full gpurun_out/prof_syrk.ncu-rep profiles/r01_syrk_ncu_full.txt [profiles/dominant_kernel_ncu.json]
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(row["Metric Unit"], 1.0)
        short = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        tot[short] += v
        cnt[short] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {sum(cnt.values())} launches, {T:.1f} ms\n")
        f.write("# share%   total_ms   launches   avg_us   kernel\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{100 * v / T:6.2f}  {v:10.3f}  {cnt[k]:8d}  {1e3 * v / cnt[k]:9.1f}   {k}\n")
    print(open(dst).read())


def full(src, dst, jdst=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {len(rows)} launches of {rows[0][idx['Kernel Name']][:80]}\n")
        for k in KEEP:
            if k in idx:
                f.write(f"{k} [{units[idx[k]]}] = {[row[idx[k]] for row in rows]}\n")
    print(open(dst).read())
    if jdst:
        def val(k, row):
            v = float(row[idx[k]].replace(",", ""))
            u = units[idx[k]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        per = [val("dram__bytes_read.sum", row) + val("dram__bytes_write.sum", row) for row in rows]
        json.dump({"kernel": rows[0][idx["Kernel Name"]][:80], "dram_bytes_per_launch": sum(per) / len(per),
                   "launches": len(rows), "source": src}, open(jdst, "w"), indent=1)


def kernel_json(src, dst, note="", flops=None):
    """One JSON per kernel (north_star: 'each kernel ships a committed ncu capture reporting achieved fp64 tensor-pipe
    utilisation and HBM GB/s against B200 peak'): averages over the captured launches of the first kernel in `src`."""
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = []
    for x in rows:
        if x[idx["Kernel Name"]] not in names:
            names.append(x[idx["Kernel Name"]])
    if len(names) > 1:                      # a capture of several kernels (the streaming kernels): one entry per kernel
        all_rows, res = rows, []
        for nm in names:
            res.append(_kernel_entry(idx, units, [x for x in all_rows if x[idx["Kernel Name"]] == nm], nm, src, note, None))
        json.dump(res, open(dst, "w"), indent=1)
        for e in res:
            print(json.dumps(e))
        return
    d = _kernel_entry(idx, units, rows, names[0], src, note, flops)
    json.dump(d, open(dst, "w"), indent=1)
    print(json.dumps(d))


def _kernel_entry(idx, units, rows, name, src, note, flops):

    def val(k):
        if k not in idx:
            alt = [h for h in idx if h.endswith("." + k)]     # e.g. "TPC.TriageCompute.sm__pipe_tensor_cycles_active..."
            if not alt:
                return None
            k = alt[0]
        vs = []
        for row in rows:
            try:
                v = float(row[idx[k]].replace(",", ""))
            except ValueError:
                return None
            u = units[idx[k]]
            v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
                  "Tbyte/s": 1e12, "Gbyte/s": 1e9}.get(u, 1.0)
            vs.append(v)
        return sum(vs) / len(vs)
    peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6549.1}
    t = val("gpu__time_duration.sum")
    dram = (val("dram__bytes_read.sum") or 0.0) + (val("dram__bytes_write.sum") or 0.0)
    d = {"kernel": re.sub(r"\(.*", "", name), "launches_captured": len(rows), "avg_duration_us": t * 1e6,
         "dram_bytes_per_launch": dram, "hbm_gbs": dram / t * 1e-9, "hbm_peak_gbs": peaks["hbm_gbs"],
         "hbm_frac": dram / t * 1e-9 / peaks["hbm_gbs"],
         "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
         "fp64_pipe_active_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
         "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "l2_hit_rate_pct": val("lts__t_sector_hit_rate.pct"),
         "registers_per_thread": val("launch__registers_per_thread"), "grid": val("launch__grid_size"),
         "block": val("launch__block_size"), "source": src, "note": note}
    if flops:
        d["flops_per_launch"] = flops
        d["tflops"] = flops / t * 1e-12
        d["fp64_pipe_peak_tflops"] = 36.9
        d["fp64_pipe_frac"] = d["tflops"] / 36.9
    return d


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "kernel":
        kernel_json(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "", float(sys.argv[5]) if len(sys.argv) > 5 else None)
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)

