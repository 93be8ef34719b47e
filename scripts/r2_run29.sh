#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run29
timeout 1200 compute-sanitizer --tool racecheck --print-limit 5000 python scripts/sanitize_probe_r2.py > ${O}_racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY" ${O}_racecheck.log | tail -3
grep -E "(Read|Write) access at" ${O}_racecheck.log | sed -E 's/.* in ([a-z_0-9]+\.cuh?:[0-9]+).*/\1/' | sort | uniq -c | sort -rn | head -30
