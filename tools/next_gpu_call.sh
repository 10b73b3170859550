#!/bin/bash
# First GPU call of the next round (one B200): validates what round 1 built after its GPU budget ran out.
#   gpurun --timeout 1500 -- 'bash tools/next_gpu_call.sh'
# Everything lands in gpurun_out/next_*.  Each step has its own timeout: phi_d_spec has never run on a GPU.
set -u
mkdir -p gpurun_out
# 1. the whole GPU suite as the driver runs it (buildhess / allreduce_dev changes included)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/next_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/next_pytest.log
# 2. the hyper-gradient sweep: parity against the per-hyper path and the oracle
OB_TEST_DSWEEP=1 timeout 300 python -m pytest tests/test_gpu_spec.py -k hyper_gradient_sweep -x -q > gpurun_out/next_dsweep_test.log 2>&1; echo "dsweep test rc=$?"; tail -5 gpurun_out/next_dsweep_test.log
# 3. what it buys: one BFGS objective evaluation with and without it
timeout 300 python tools/dsweep_bench.py --config c3 > gpurun_out/next_dsweep_c3.json 2> gpurun_out/next_dsweep_c3.err; echo "dsweep c3 rc=$?"; cat gpurun_out/next_dsweep_c3.json
timeout 600 python tools/dsweep_bench.py --config c4share > gpurun_out/next_dsweep_c4.json 2> gpurun_out/next_dsweep_c4.err; echo "dsweep c4 rc=$?"; cat gpurun_out/next_dsweep_c4.json
# 4. ncu of the new kernel (only after the runs above exited 0)
OB_DSWEEP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:phi_d_spec -c 2 -o gpurun_out/next_phi_d \
  python tools/dsweep_bench.py --config c3 > gpurun_out/next_ncu.log 2>&1; echo "ncu rc=$?"
