#!/bin/bash
# A/B of library builds: tools/probes/ab_lib.sh A B ...   (arap_flow_b200/libarapb200_<name>.so)
cd $GRAFT_REPO_ROOT
cp arap_flow_b200/libarapb200.so /tmp/orig.so
for v in "$@"; do
  cp arap_flow_b200/libarapb200_$v.so arap_flow_b200/libarapb200.so
  for i in 1 2; do timeout 400 python bench.py --workload C1 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['e2e']['value'])"; done
done
cp /tmp/orig.so arap_flow_b200/libarapb200.so
