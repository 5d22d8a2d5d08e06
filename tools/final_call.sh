#!/bin/bash
# One GPU-box call at the end of a round: smoke, the bench line, the ncu launch list of the same code, the GPU suite.
# Everything lands in gpurun_out/ step by step, so a call that is cut short still leaves what it finished.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
( time timeout 600 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" >> gpurun_out/bench_n1.err
( time timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --skip-cpu --steps 2 --no-extras ) > gpurun_out/ncu.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/smoke.log; tail -5 gpurun_out/pytest_gpu.log; tail -c 600 gpurun_out/bench_n1.json
