#!/bin/bash
# A/B of the GEMM work-item scheduler (AGNN_GEMM_SCHED) and the GRU kernel choice (AGNN_GRU_TC) on the headline step.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
( timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_f16x3_gpu.py tests/test_fused_gpu.py tests/test_gru_gpu.py tests/test_train.py -m gpu -x -q ) > gpurun_out/ab_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/ab_pytest.log
for cfg in "static 2" "dynamic 2" "dynamic 1" "static 1" "dynamic 2"; do
  set -- $cfg
  AGNN_GEMM_SCHED=$1 AGNN_GRU_TC=$2 timeout 200 python bench.py --skip-cpu --no-extras --steps 20 2> gpurun_out/ab_$1_$2.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 gru_tc=$2', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['loss'])" >> gpurun_out/ab_sched.txt 2>&1
done
tail -3 gpurun_out/ab_pytest.log; cat gpurun_out/ab_sched.txt
