#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run43
for tool in memcheck racecheck synccheck; do
timeout 500 compute-sanitizer --tool $tool --print-limit 2000 python scripts/sanitize_probe_r2b.py > ${O}_$tool.log 2>&1; echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|probe done" ${O}_$tool.log | tail -3
done
grep -E "(Read|Write) access at|Race reported" gpurun_out/r2_run43_racecheck.log | sed -E 's/.* in ([a-z_0-9]+\.cuh?:[0-9]+).*/\1/' | sort | uniq -c | sort -rn | head -20
