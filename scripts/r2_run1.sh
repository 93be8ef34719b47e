#!/bin/bash
# round 2, GPU run 1: full GPU test suite after the scratch-context refactor, then builder probes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_run1_pytest.log
tail -5 gpurun_out/r2_run1_pytest.log
for cfg in "--n 20000 --dim 64 --M 16 --efc 100 --parity 256 --oracle" \
           "--n 100000 --dim 128 --M 16 --efc 200 --parity 256 --oracle" \
           "--n 100000 --dim 768 --M 32 --efc 200 --parity 256" \
           "--n 1000000 --dim 768 --M 32 --efc 200 --parity 256" \
           "--n 200000 --dim 1024 --M 32 --efc 200 --dtype bin1" \
           "--n 200000 --dim 768 --M 32 --efc 200 --dtype f16"; do
  VELES_BUILD_VERBOSE=1 timeout 900 python scripts/probe_build.py $cfg >> gpurun_out/r2_run1_build.jsonl 2>> gpurun_out/r2_run1_build.err
  echo "exit $? for $cfg" >> gpurun_out/r2_run1_build.err
done
cat gpurun_out/r2_run1_build.jsonl
tail -20 gpurun_out/r2_run1_build.err
