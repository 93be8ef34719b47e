#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run46
timeout 600 python bench.py --config c2sq8 --steps 10 > ${O}_c2sq8.json 2> ${O}_c2sq8.err; echo "c2sq8 exit $?"
timeout 600 python bench.py --config c4 --n 4000000 --steps 10 --cpu-seconds 5 > ${O}_c4_4M.json 2> ${O}_c4_4M.err; echo "c4 exit $?"
python - <<'PY'
import json
for f in ("c2sq8","c4_4M"):
    try:
        d=json.loads(open(f"gpurun_out/r2_run46_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), d.get("roofline",{}).get("frac"), d.get("e2e",{}).get("value"), d.get("clocks",{}).get("reasons"), d.get("gpu_launches"), d.get("parity"), [v for k,v in d.items() if k.startswith("recall")])
    except Exception as e:
        print(f, "ERR", e)
PY
