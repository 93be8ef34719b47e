#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run21
timeout 300 python -m pytest tests/test_gpu_bm25_fusion.py -m gpu -x -q > ${O}_pytest_bm25.log 2>&1; echo "bm25 pytest exit $?" >> ${O}_pytest_bm25.log; tail -6 ${O}_pytest_bm25.log
timeout 600 python scripts/probe_bm25.py > ${O}_bm25.jsonl 2> ${O}_bm25.err; echo "bm25 exit $?"; cat ${O}_bm25.jsonl; tail -3 ${O}_bm25.err
timeout 900 ncu --clock-control none --set full --import-source on -k regex:bm25_sub_kernel --launch-skip 2 --launch-count 1 -f -o ${O}_prof_bm25 python scripts/probe_bm25.py default > ${O}_prof_bm25.log 2>&1; echo "bm25 ncu exit $?"
