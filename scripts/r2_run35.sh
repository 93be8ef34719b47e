#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run35
VELES_BF_DEBUG_TIMING=1 timeout 300 python scripts/probe_bfq1.py > ${O}_q1_dbg.jsonl 2> ${O}_q1_dbg.err; echo "exit $?"; grep "bf timing" ${O}_q1_dbg.err | sed -n 25,32p
