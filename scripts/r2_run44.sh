#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run44
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -3 ${O}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 ${O}_smoke.log
timeout 900 python bench.py --impl reference > ${O}_ref_c2.json 2> ${O}_ref_c2.err; echo "ref exit $?"; tail -c 300 ${O}_ref_c2.json
timeout 900 python bench.py > ${O}_c2.json 2> ${O}_c2.err; echo "c2 exit $?"; tail -c 600 ${O}_c2.json
timeout 900 python bench.py --config c5 --steps 10 > ${O}_c5.json 2> ${O}_c5.err; echo "c5 exit $?"
timeout 900 python bench.py --config c1 --steps 10 > ${O}_c1.json 2> ${O}_c1.err; echo "c1 exit $?"
python - <<'PY'
import json
for f in ("ref_c2","c2","c5","c1"):
    try:
        d=json.loads(open(f"gpurun_out/r2_run44_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), d.get("roofline",{}).get("frac"), d.get("e2e",{}).get("value"), d.get("clocks",{}).get("reasons"), d.get("gpu_launches"), d.get("relaxed_mode"), d.get("parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
