#!/bin/bash
# round 2, GPU run 2: new bench.py on every config (scaled where noted), reference arm from frozen files, tcgen05 GEMM
mkdir -p gpurun_out
O=gpurun_out/r2_run2
( free -g; nproc; df -h /tmp | tail -1; nvidia-smi --query-gpu=name,memory.total --format=csv ) > ${O}_box.txt 2>&1
run() { name=$1; shift; timeout 1500 python bench.py "$@" > ${O}_$name.json 2> ${O}_$name.err; echo "$name exit $?"; tail -c 1800 ${O}_$name.json; echo; tail -3 ${O}_$name.err; }
run ref_c2 --impl reference --config c2
run c2 --config c2
run c2sq8 --config c2sq8
run c1 --config c1
run c5 --config c5 --steps 5
run c3s --config c3 --n 2000000 --steps 10
run c4s --config c4 --n 4000000 --steps 10
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -15 ${O}_pytest.log
timeout 300 python scripts/probe_gemm.py --n 1000000 --dim 768 --nq 1024 --k 10 > ${O}_gemm.jsonl 2> ${O}_gemm.err; echo "gemm exit $?"; cat ${O}_gemm.jsonl; tail -3 ${O}_gemm.err
