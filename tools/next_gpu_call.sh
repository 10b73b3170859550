#!/bin/bash
# Measurement pass of a round on ONE B200 (what profiles/r02_* were made with):
#   gpurun --timeout 1500 -- 'bash tools/next_gpu_call.sh r02'
# Everything lands in gpurun_out/<tag>_*.  Every step has its own timeout; nothing timed under ncu is a bench value.
set -u
T=${1:-r02}
O=gpurun_out
mkdir -p $O
# 1. the whole GPU suite as the driver runs it
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${T}_pytest_gpu.log
# 2. the contract line (C3) and the other configs
timeout 600 python bench.py > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err; echo "bench rc=$?"; cat $O/${T}_bench_1gpu.json
for c in c2 c4share c5 build; do
  timeout 600 python bench.py --config $c > $O/${T}_bench_$c.json 2> $O/${T}_bench_$c.err; echo "bench $c rc=$?"; cat $O/${T}_bench_$c.json
done
# 3. launch list of the bench command (share of each kernel in the step).  OB_OVERLAP=0: ncu serialises kernels and copies,
#    so the streamed-input kernel of the e2e leg would wait for rows that cannot arrive
OB_OVERLAP=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_launches.log 2>&1; echo "launch list rc=$?"
# 4. full captures: the two products of the step, then the multi-RHS tensor-core kernels of C5
OB_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:phi_._spec -s 6 -c 2 -o $O/${T}_spec -f \
  python bench.py --steps 2 --warmup 3 --no-optcg --no-cpu-baseline > $O/${T}_ncu_spec.log 2>&1; echo "ncu spec rc=$?"
OB_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:phi_.m_spec -s 2 -c 2 -o $O/${T}_mat -f \
  python bench.py --config c5 --steps 2 --warmup 2 --no-cpu-baseline > $O/${T}_ncu_mat.log 2>&1; echo "ncu mat rc=$?"
ls -la $O | tail -20
