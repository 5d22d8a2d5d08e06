#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for cfg in "1 0" "1 3" "1 2"; do
  set -- $cfg
  AGNN_SIDE_PRIORITY=$1 AGNN_GRU_TC=$2 timeout 200 python bench.py --skip-cpu --no-extras --steps 20 2> gpurun_out/abp2_$1_$2.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prio=$1 gru_tc=$2', d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])" >> gpurun_out/ab_prio2.txt 2>&1
done
for prio in 0 1; do
  for which in config4 config3; do
    AGNN_SIDE_PRIORITY=$prio timeout 200 python tools/config_probe.py $which 2> gpurun_out/abp2_${which}_$prio.err | \
      python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
def walk(o,pre=''):
    if isinstance(o,dict):
        for k,v in o.items():
            if isinstance(v,(dict,)): walk(v,pre+k+'.')
            elif 'ms_per_step' in k: print('prio=$prio $which', pre+k, v)
walk(d)" >> gpurun_out/ab_prio2.txt 2>&1
  done
done
cat gpurun_out/ab_prio2.txt
