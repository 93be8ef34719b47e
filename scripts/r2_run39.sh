#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run39
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -8 ${O}_pytest.log
timeout 300 ncu --clock-control none --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file ${O}_bf_ncu.csv python scripts/probe_bf_ncu.py > ${O}_bf_ncu.jsonl 2> ${O}_bf_ncu.err; echo "ncu exit $?"; grep -v "^==" ${O}_bf_ncu.csv | cut -d, -f5,13- | head -20
timeout 300 python scripts/probe_bf2.py > ${O}_bf2.jsonl 2> ${O}_bf2.err; echo "bf2 exit $?"; cat ${O}_bf2.jsonl
for d in 3 4; do
timeout 300 python bench.py --config c2 --steps 40 --no-cpu-baseline --pipe-depth $d > ${O}_c2_d$d.json 2> ${O}_c2_d$d.err; echo "c2 depth $d exit $?"; python -c "
import json;l=json.load(open('${O}_c2_d$d.json'));print(l['value'],l['roofline']['frac'],l['e2e'])"
done
