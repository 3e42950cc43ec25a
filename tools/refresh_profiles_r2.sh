#!/bin/bash
# Round-2 measurement pass on the GPU box (through gpurun; everything lands in gpurun_out/, summaries are then written to
# profiles/ by tools/summarize_ncu_r2.py).  Order: tests -> smoke -> benches (no profiler) -> ncu captures of commands that
# have already exited 0 without ncu.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_tests.log 2>&1; tail -3 gpurun_out/r2_tests.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
for w in C1 C1s C0 C2 C3 C4; do
  extra="--no-cpu-baseline"; [ "$w" = "C1" ] && extra=""
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 $extra > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err || echo "bench $w failed"
  python -c "import json;d=json.load(open('gpurun_out/r2_bench_$w.json'));print('$w',round(d['value'],4),round(d['e2e']['value'],4),round(d['roofline']['frac'],3),d['run']['launch'])"
done
timeout 600 python bench.py --path opt_h --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_opt_h.json 2> gpurun_out/r2_bench_opt_h.err
python -c "import json;d=json.load(open('gpurun_out/r2_bench_opt_h.json'));print('opt_h',round(d['value'],4),d['run']['opt_h_seconds_last_image'])"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python -c "import json;d=json.load(open('gpurun_out/r2_bench_reference.json'));print('reference',d['value'],d['ms_per_step'],d['extrapolated'],d['cpu_baseline']['cores'])"
# launch list of the default bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launch_list.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1 || echo "ncu launch list failed"
# the dominant kernel, full set, on the default bench's launch shape (NP co-resident C1 problems) at a reduced schedule
NP=${NP:-4}
timeout 200 python tools/ncu_target.py resident C1 50 $NP > gpurun_out/r2_ncu_target_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_resident --launch-skip 1 --launch-count 1 -f -o gpurun_out/r2_resident_b$NP python tools/ncu_target.py resident C1 50 $NP > gpurun_out/r2_ncu_full.log 2>&1 || echo "ncu full failed"
# DRAM traffic of one FULL-schedule launch of NP co-resident problems
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_resident --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/r2_traffic_resident$NP.csv python bench.py --steps 1 --warmup 1 --batch $NP --no-cpu-baseline > gpurun_out/r2_ncu_traffic$NP.log 2>&1 || echo "traffic capture failed"
tail -2 gpurun_out/r2_traffic_resident$NP.csv
python tools/profile_resident.py C1 > gpurun_out/r2_prof_c1.log 2>&1; cat gpurun_out/r2_prof_c1.log
