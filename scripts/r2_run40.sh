#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run40
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -8 ${O}_pytest.log
timeout 300 python scripts/probe_gemm.py > ${O}_gemm_new.jsonl 2> ${O}_gemm_new.err; echo "new exit $?"; cat ${O}_gemm_new.jsonl
VELES_TC_OLD_TAIL=1 timeout 300 python scripts/probe_gemm.py --skip-exact > ${O}_gemm_old.jsonl 2> ${O}_gemm_old.err; echo "old exit $?"; cat ${O}_gemm_old.jsonl
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file ${O}_gemm_launches.csv python scripts/probe_gemm.py --skip-exact > /dev/null 2> ${O}_gemm_ncu.err; echo "ncu exit $?"; grep -v "^==" ${O}_gemm_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
timeout 300 python bench.py --config c1 --steps 20 --no-cpu-baseline > ${O}_c1.json 2> ${O}_c1.err; echo "c1 exit $?"; python -c "
import json;l=json.load(open('${O}_c1.json'));print(l['value'],l['ms_per_step'])"
