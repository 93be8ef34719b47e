#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run9
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -6 ${O}_pytest.log
timeout 300 python scripts/probe_bf2.py > ${O}_bf.jsonl 2> ${O}_bf.err; echo "bf exit $?"; cat ${O}_bf.jsonl; tail -3 ${O}_bf.err
VELES_BF_NO_FUSE=1 timeout 300 python scripts/probe_bf2.py > ${O}_bf_nofuse.jsonl 2> ${O}_bf_nofuse.err; head -2 ${O}_bf_nofuse.jsonl
timeout 600 ncu --clock-control none --set full -k regex:bf_scan_kernel --launch-skip 3 --launch-count 1 -f -o ${O}_prof_bfscan python scripts/probe_bf2.py > ${O}_prof_bfscan.log 2>&1; echo "bf ncu exit $?"
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 ${O}_smoke.log
