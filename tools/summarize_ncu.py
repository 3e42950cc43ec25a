#!/usr/bin/env python
"""Turn the .ncu-rep / launch-list files in gpurun_out/ into the text summaries committed under profiles/."""
import collections, csv, io, subprocess, sys, os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

def launch_list(path, out, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        k = r[ik].split("(")[0].replace("arapb200::<unnamed>::", "").replace("void ", "")
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"# {title}", "# ncu --metrics gpu__time_duration.sum --clock-control none ; cold-cache, serialised: compare SHARES",
             f"# total {tot / 1e3:.1f} ms over {sum(v[0] for v in agg.values())} launches", "kernel,launches,total_us,share"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"{k},{n},{t:.1f},{t / tot:.5f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:8]))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]

def full(rep, notes):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"## {os.path.basename(rep)} -- {notes}"]
    for r in rows[2:]:
        lines.append("- " + r[idx["Kernel Name"]].replace("arapb200::<unnamed>::", "")[:70])
        for w in WANT:
            if w in idx:
                lines.append(f"    {w} [{units[idx[w]]}] = {r[idx[w]]}")
    return lines

def stalls(rep, warps, iters):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    hdr, data = rows[starts[-2] + 1], rows[starts[-2] + 2:starts[-1]]   # the last captured launch
    cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in data:
        for i, h in cols:
            try: tot[h] += int(r[i])
            except ValueError: pass
    s = sum(tot.values())
    iex = hdr.index("Instructions Executed")
    ninst = sum(int(r[iex]) for r in data)
    lines = ["", f"warp stall sampling (all samples); {ninst} warp-instructions = {ninst / warps / iters:.0f} per warp per PCG-iteration-equivalent"]
    for h, v in tot.most_common():
        if v: lines.append(f"  {h:28s} {v:7d}  {100 * v / s:5.1f} %")
    return lines

if __name__ == "__main__":
    launch_list(os.path.join(G, "bench_launch_list.csv"), os.path.join(P, "r1_bench_launch_list.csv"),
                "ncu launch list: python bench.py --steps 1 --warmup 1 --no-cpu-baseline (round 1; 2 steps x 9 pairs, C1)")
    L = full(os.path.join(G, "r1_resident_b3_final.ncu-rep"), "tools/ncu_target.py resident C1 50 3: 3 co-resident 854x480 problems (the default bench's launch shape, 168-register variant), 1x1x50 PCG iterations")
    L += stalls(os.path.join(G, "r1_resident_b3_final.ncu-rep"), 584 * 3, 52)
    L += [""] + full(os.path.join(G, "r1_resident_b4_final.ncu-rep"), "tools/ncu_target.py resident C1 50 4: 4 co-resident problems (128-register variant)")
    L += stalls(os.path.join(G, "r1_resident_b4_final.ncu-rep"), 584 * 4, 52)
    L += [""] + full(os.path.join(G, "r1_stream_c4_final.ncu-rep"), "tools/ncu_target.py stream C4 8 1: 1920x1080 (1 378 443 active px) through the streaming back-end "
                     "(tile-interleaved layout; warm-cache per-kernel times of the final kernels: r1_stream_kernel_times.txt)")
    open(os.path.join(P, "r1_ncu_full_summary.txt"), "w").write("\n".join(L) + "\n")
    print("\n".join(L[:40]))
