#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run42
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "tensor_core" > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -8 ${O}_pytest.log
for d in 64 32 16; do
VELES_TC_SAMPLE_DIV=$d timeout 300 python scripts/probe_gemm.py > ${O}_gemm_div$d.jsonl 2> ${O}_gemm_div$d.err; echo "div $d exit $?"; cat ${O}_gemm_div$d.jsonl
done
VELES_TC_REGTOPK=1 timeout 300 python scripts/probe_gemm.py --skip-exact > ${O}_gemm_regtopk.jsonl 2> ${O}_gemm_regtopk.err; cat ${O}_gemm_regtopk.jsonl
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file ${O}_gemm_launches.csv python scripts/probe_gemm.py --skip-exact > /dev/null 2> ${O}_gemm_ncu.err; echo "ncu exit $?"; grep -v "^==" ${O}_gemm_launches.csv | awk -F'","' '{print substr($5,1,60), $NF}' | tail -6
timeout 300 python scripts/probe_gemm.py --n 4000000 --k 100 --over 1 --skip-exact > ${O}_gemm_4M_k100.jsonl 2>${O}_gemm_4M.err; cat ${O}_gemm_4M_k100.jsonl
timeout 300 python scripts/probe_gemm.py --nq 256 --skip-exact > ${O}_gemm_q256.jsonl 2>${O}_gemm_q256.err; cat ${O}_gemm_q256.jsonl
