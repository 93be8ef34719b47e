#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run36
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -5 ${O}_pytest.log
timeout 300 python scripts/probe_bfq1.py > ${O}_q1.jsonl 2> ${O}_q1.err; echo "exit $?"; cat ${O}_q1.jsonl
VELES_BF_DEBUG_TIMING=1 timeout 300 python scripts/probe_bfq1.py > ${O}_q1_dbg.jsonl 2> ${O}_q1_dbg.err; echo "exit $?"; grep "bf timing" ${O}_q1_dbg.err | sed -n 25,32p
