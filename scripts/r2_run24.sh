#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run24
timeout 900 ncu --clock-control none --set full --import-source on -k regex:bf_scan_kernel --launch-skip 5 --launch-count 1 -f -o ${O}_prof_bfq1 python scripts/probe_bf2.py > ${O}_prof_bfq1.log 2>&1; echo "ncu exit $?"
