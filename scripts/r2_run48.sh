#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run48
timeout 600 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -3 ${O}_pytest.log
timeout 200 python bench.py --config c5 --steps 5 --no-cpu-baseline --hybrid-three-calls > ${O}_c5_three.json 2> ${O}_c5_three.err; echo "c5 three exit $?"; python -c "
import json;l=json.load(open('${O}_c5_three.json'));print(l['value'],l['ms_per_step'],l['e2e'])"; tail -2 ${O}_c5_three.err
