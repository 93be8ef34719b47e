#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run23
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -4 ${O}_pytest.log
timeout 600 python scripts/probe_bf2.py > ${O}_bf.jsonl 2> ${O}_bf.err; echo "bf exit $?"; cat ${O}_bf.jsonl | cut -c1-150; tail -3 ${O}_bf.err
VELES_BF_NO_DEEP=1 timeout 600 python scripts/probe_bf2.py > ${O}_bf_nodeep.jsonl 2> ${O}_bf_nodeep.err; echo "bf nodeep exit $?"; head -3 ${O}_bf_nodeep.jsonl | cut -c1-150
