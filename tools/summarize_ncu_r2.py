#!/usr/bin/env python
"""Round 2: gpurun_out/r2_* (tools/refresh_profiles_r2.sh) -> the summaries committed under profiles/."""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import summarize_ncu as S1  # noqa: E402  (launch_list, full, stalls)

G, P = S1.G, S1.P


def main():
    ll = os.path.join(G, "r2_bench_launch_list.csv")
    if os.path.exists(ll):
        S1.launch_list(ll, os.path.join(P, "r2_bench_launch_list.csv"),
                       "ncu launch list: python bench.py --steps 1 --warmup 1 --no-cpu-baseline (round 2; 2 steps x 8 pairs, C1)")
    NP = int(os.environ.get("NP", "4"))         # co-resident problems of the default bench's launches
    regs = {3: 168, 4: 128}.get(NP, 0)
    rep = os.path.join(G, f"r2_resident_b{NP}.ncu-rep")
    if os.path.exists(rep):
        L = S1.full(rep, f"tools/ncu_target.py resident C1 50 {NP}: {NP} co-resident 854x480 problems (the default bench's launch shape, "
                         f"{regs}-register variant), 1x1x50 PCG iterations; ncu --set full --clock-control none --import-source on")
        L += S1.stalls(rep, 584 * NP, 52)
        open(os.path.join(P, "r2_ncu_full_summary.txt"), "w").write("\n".join(L) + "\n")
        print("\n".join(L[:30]))
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, r = rows[0], rows[-1]
        want = ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
        d = {"source": f"profiles/r2_ncu_full_summary.txt (gpurun_out/r2_resident_b{NP}.ncu-rep: {NP} co-resident C1 problems, {regs}-register variant)"}
        for w in want:
            if w in hdr:
                try:
                    d[w] = float(r[hdr.index(w)].replace(",", ""))
                except ValueError:
                    d[w] = r[hdr.index(w)]
        json.dump(d, open(os.path.join(P, "r2_ncu_resident.json"), "w"), indent=1)
    tr = os.path.join(G, f"r2_traffic_resident{NP}.csv")
    if os.path.exists(tr):
        rows = [r for r in csv.reader(open(tr)) if len(r) > 10]
        hdr = rows[0]
        vals = {r[hdr.index("Metric Name")]: float(r[hdr.index("Metric Value")].replace(",", "")) for r in rows[1:]}
        k = rows[1][hdr.index("Kernel Name")].split("::")[-1]
        alg = 135213 * (156.0 * 60800 + 132.0 * 152) * NP
        prev = {}
        try:
            prev = json.load(open(os.path.join(P, "r2_traffic.json")))["launches"]
        except Exception:
            pass
        out = {"what": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_resident "
                       "--launch-skip 1 --launch-count 1 on: python bench.py --steps 1 --warmup 1 --batch N --no-cpu-baseline; one full 19x8x400 "
                       "launch of N co-resident C1 problems (round 2)",
               "launches": {str(NP): {"kernel": f"{k}, grid {rows[1][hdr.index('Grid Size')]}, block {rows[1][hdr.index('Block Size')]}",
                                  "problems_per_launch": NP, "dram_bytes_read": int(vals["dram__bytes_read.sum"]),
                                  "dram_bytes_write": int(vals["dram__bytes_write.sum"]),
                                  "dram_bytes_per_launch": int(vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]),
                                  "duration_ms_under_ncu": vals["gpu__time_duration.sum"] / 1e6, "algorithmic_bytes_per_launch": alg}}}
        for k2, v2 in prev.items():                 # keep the captures of the other launch shapes
            out["launches"].setdefault(k2, v2)
        json.dump(out, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
        shutil.copy(tr, os.path.join(P, f"r2_traffic_resident{NP}_ncu.csv"))
        print(json.dumps(out["launches"], indent=1))
    for w in ("C1", "C1s", "C0", "C2", "C3", "C4", "opt_h", "reference"):
        src = os.path.join(G, f"r2_bench_{w}.json")
        if os.path.exists(src) and os.path.getsize(src) > 10:
            shutil.copy(src, os.path.join(P, f"r2_bench_{w}.json"))
    for f in ("r2_prof_c1.log",):
        if os.path.exists(os.path.join(G, f)):
            shutil.copy(os.path.join(G, f), os.path.join(P, "r2_resident_cycle_accounting.txt"))


if __name__ == "__main__":
    main()
