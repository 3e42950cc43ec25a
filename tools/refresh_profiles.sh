#!/bin/bash
# One GPU-box pass that re-measures everything committed under profiles/ (run through gpurun; results land in gpurun_out/).
mkdir -p gpurun_out
for w in C1 C0 C2 C3 C4; do
  extra=""; [ "$w" != "C1" ] && extra="--no-cpu-baseline"
  timeout 500 python bench.py --workload $w $extra > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || echo "bench $w failed"
  python -c "import json;d=json.load(open('gpurun_out/bench_$w.json'));print('$w',round(d['value'],4),round(d['e2e']['value'],4),round(d['roofline']['frac'],3),d['config']['backend'])"
done
timeout 300 python tools/rtol_sweep.py C1 4 > gpurun_out/rtol_C1.log 2>&1; tail -3 gpurun_out/rtol_C1.log
timeout 300 python tools/rtol_sweep.py cat512 4 > gpurun_out/rtol_cat512.log 2>&1; tail -3 gpurun_out/rtol_cat512.log
# launch list of the default bench command (only after it exited 0 without ncu above)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/bench_launch_list.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1 || echo "ncu launch list failed"
# streaming kernels, warm-cache per-kernel times
tools/kernel_times.sh stream C4 8 1 > gpurun_out/stream_kernel_times.txt 2>&1; grep -E "k_step|k_init" gpurun_out/stream_kernel_times.txt
# DRAM traffic of one full-schedule resident launch (4 co-resident C1 problems)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_resident --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/traffic_resident.csv python bench.py --steps 1 --warmup 1 --batch 4 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1 || echo "traffic capture failed"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_resident --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/traffic_resident3.csv python bench.py --steps 1 --warmup 1 --batch 3 --no-cpu-baseline > gpurun_out/ncu_traffic3.log 2>&1 || echo "traffic capture (3) failed"
tail -2 gpurun_out/traffic_resident.csv; tail -2 gpurun_out/traffic_resident3.csv   # then: python tools/update_traffic.py
timeout 300 python tools/cli_throughput.py 64 > gpurun_out/cli_throughput.txt 2>&1; cat gpurun_out/cli_throughput.txt
