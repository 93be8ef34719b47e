#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run38
timeout 300 python -m pytest tests/test_gpu_bm25_fusion.py -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -15 ${O}_pytest.log
timeout 400 python bench.py --config c5 --steps 20 --no-cpu-baseline > ${O}_c5.json 2> ${O}_c5.err; echo "c5 exit $?"; cat ${O}_c5.json; tail -5 ${O}_c5.err
timeout 300 ncu --clock-control none --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file ${O}_bf_ncu.csv python scripts/probe_bf_ncu.py > ${O}_bf_ncu.jsonl 2> ${O}_bf_ncu.err; echo "ncu exit $?"; cat ${O}_bf_ncu.jsonl; tail -3 ${O}_bf_ncu.err
