#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run22
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -4 ${O}_pytest.log
timeout 600 python scripts/probe_bm25.py all > ${O}_bm25.jsonl 2> ${O}_bm25.err; echo "bm25 exit $?"; cat ${O}_bm25.jsonl | cut -c1-120; tail -3 ${O}_bm25.err
timeout 900 ncu --clock-control none --set full --import-source on -k regex:bm25_sub_kernel --launch-skip 2 --launch-count 1 -f -o ${O}_prof_bm25 python scripts/probe_bm25.py default > ${O}_prof_bm25.log 2>&1; echo "bm25 ncu exit $?"
timeout 900 python bench.py --config c5 --steps 5 > ${O}_c5.json 2> ${O}_c5.err; echo "c5 exit $?"; tail -c 300 ${O}_c5.json
