#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run28
timeout 200 python scripts/sanitize_probe_r2.py > ${O}_plain.log 2>&1; echo "plain exit $?"; tail -2 ${O}_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_probe_r2.py > ${O}_${tool}.log 2>&1; echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" ${O}_${tool}.log | tail -3
done
