#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run47
timeout 300 python -m pytest tests/test_gpu_bm25_fusion.py -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -15 ${O}_pytest.log
