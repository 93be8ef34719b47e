#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run31
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "binary or bin" > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -5 ${O}_pytest.log
timeout 900 python bench.py --config c4 --n 4000000 --steps 10 --no-cpu-baseline > ${O}_c4s_r4.json 2> ${O}_c4s_r4.err; echo "c4s r4 exit $?"; tail -c 200 ${O}_c4s_r4.err
VELES_SEARCH_BIN1_R4=0 timeout 900 python bench.py --config c4 --n 4000000 --steps 10 --no-cpu-baseline > ${O}_c4s_r8.json 2> ${O}_c4s_r8.err; echo "c4s r8 exit $?"
python - <<'PY'
import json
for f in ("r4","r8"):
    d=json.loads(open(f"gpurun_out/r2_run31_c4s_{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("recall"), d["parity"], d["e2e"]["value"])
PY
