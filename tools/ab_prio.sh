#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for cfg in "0 2" "1 2" "1 1" "0 2" "1 2"; do
  set -- $cfg
  AGNN_SIDE_PRIORITY=$1 AGNN_GRU_TC=$2 timeout 200 python bench.py --skip-cpu --no-extras --steps 20 2> gpurun_out/abp_$1_$2.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prio=$1 gru_tc=$2', d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])" >> gpurun_out/ab_prio.txt 2>&1
done
cat gpurun_out/ab_prio.txt
