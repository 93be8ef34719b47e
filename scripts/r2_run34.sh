#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run34
timeout 300 python -m pytest tests/test_gpu_bm25_fusion.py -m gpu -x -q > ${O}_pytest_bm25.log 2>&1; echo "bm25 pytest exit $?" >> ${O}_pytest_bm25.log; tail -3 ${O}_pytest_bm25.log
timeout 600 python scripts/probe_bm25.py > ${O}_bm25.jsonl 2> ${O}_bm25.err; echo "bm25 exit $?"; cut -c1-120 ${O}_bm25.jsonl; tail -3 ${O}_bm25.err
