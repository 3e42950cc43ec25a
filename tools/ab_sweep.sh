#!/bin/bash
# A/B of library variants on the GPU box: C1 bench value + solo / co-resident iteration times.  Usage: tools/ab_sweep.sh name1 name2 ...
# ("default" = arap_flow_b200/libarapb200.so, anything else = arap_flow_b200/variants/libarapb200_<name>.so)
for v in "$@"; do
  if [ "$v" = default ]; then unset ARAPB200_LIB; else export ARAPB200_LIB=$PWD/arap_flow_b200/variants/libarapb200_$v.so; fi
  val=$(python bench.py --steps ${STEPS:-4} --warmup 2 --no-cpu-baseline ${BENCH_ARGS} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3f pairs/s (e2e %.3f), single-problem GN step %s ms' % (d['value'], d['e2e']['value'], d.get('ms_per_gn_step_single_problem')))")
  echo "$v: $val"
done
