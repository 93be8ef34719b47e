#!/bin/bash
# round 2, GPU run 8 (1 GPU): after the branch-free GEMM epilogue and the host-overhead caches; final single-GPU lines
mkdir -p gpurun_out
O=gpurun_out/r2_run8
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -12 ${O}_pytest.log
timeout 300 python scripts/probe_bf2.py > ${O}_bf.jsonl 2> ${O}_bf.err; echo "bf exit $?"; cat ${O}_bf.jsonl; tail -3 ${O}_bf.err
timeout 300 python scripts/probe_gemm.py --n 1000000 --dim 768 --nq 1024 --k 10 > ${O}_gemm.jsonl 2> ${O}_gemm.err; echo "gemm exit $?"; cat ${O}_gemm.jsonl; tail -3 ${O}_gemm.err
timeout 300 python scripts/probe_gemm.py --n 1000000 --dim 768 --nq 256 --k 10 --skip-exact >> ${O}_gemm.jsonl 2>> ${O}_gemm.err; tail -1 ${O}_gemm.jsonl
timeout 300 python scripts/probe_gemm.py --n 4000000 --dim 768 --nq 1024 --k 100 --skip-exact >> ${O}_gemm.jsonl 2>> ${O}_gemm.err; tail -1 ${O}_gemm.jsonl
timeout 600 python scripts/probe_bm25.py > ${O}_bm25.jsonl 2> ${O}_bm25.err; echo "bm25 exit $?"; cat ${O}_bm25.jsonl; tail -3 ${O}_bm25.err
timeout 900 ncu --clock-control none --set full --import-source on -k regex:gemm_tc_kernel --launch-skip 5 --launch-count 1 -f -o ${O}_prof_gemm python scripts/probe_gemm.py --n 1000000 --skip-exact > ${O}_prof_gemm.log 2>&1; echo "gemm ncu exit $?"
timeout 900 ncu --clock-control none --set full --import-source on -k regex:bf_scan_kernel --launch-skip 3 --launch-count 1 -f -o ${O}_prof_bfscan python scripts/probe_bf2.py > ${O}_prof_bfscan.log 2>&1; echo "bf ncu exit $?"
run() { name=$1; shift; timeout 1500 python bench.py "$@" > ${O}_$name.json 2> ${O}_$name.err; echo "$name exit $?"; tail -c 1200 ${O}_$name.json; echo; tail -3 ${O}_$name.err; }
run ref_c2 --impl reference --config c2
run c2 --config c2
run c2sq8 --config c2sq8
run c5 --config c5 --steps 5
run c1 --config c1
