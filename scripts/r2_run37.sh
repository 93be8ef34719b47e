#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run37
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -5 ${O}_pytest.log
timeout 300 python scripts/probe_bfq1.py > ${O}_q1.jsonl 2> ${O}_q1.err; echo "exit $?"; cat ${O}_q1.jsonl
