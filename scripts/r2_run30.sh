#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run30
timeout 1200 ncu --clock-control none --set full --import-source on --profile-from-start off -k regex:hnsw_search_kernel --launch-count 1 -f -o ${O}_prof_c4 python bench.py --config c4 --n 4000000 --steps 2 --no-cpu-baseline > ${O}_prof_c4.log 2>&1; echo "ncu exit $?"; tail -c 400 ${O}_prof_c4.log
