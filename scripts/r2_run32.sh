#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run32
timeout 1500 ncu --clock-control none --set full --import-source on --profile-from-start off -k regex:hnsw_search_kernel --launch-count 1 -f -o ${O}_prof_c3 python bench.py --config c3 --steps 2 --no-cpu-baseline > ${O}_prof_c3.log 2>&1; echo "ncu exit $?"; tail -c 300 ${O}_prof_c3.log
