#!/bin/bash
# round 2, GPU run 3: fused brute force, ef sweep for the binary config, C3 and C4 at their stated sizes
mkdir -p gpurun_out
O=gpurun_out/r2_run3
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -15 ${O}_pytest.log
timeout 300 python scripts/probe_bf2.py > ${O}_bf.jsonl 2> ${O}_bf.err; echo "bf exit $?"; cat ${O}_bf.jsonl; tail -3 ${O}_bf.err
timeout 300 python scripts/probe_gemm.py --n 1000000 --dim 768 --nq 1024 --k 10 > ${O}_gemm.jsonl 2> ${O}_gemm.err; echo "gemm exit $?"; cat ${O}_gemm.jsonl
timeout 600 python scripts/probe_ef.py --config c4 --n 4000000 --efs 64,128,256,512,1024 > ${O}_ef_c4.jsonl 2> ${O}_ef_c4.err; echo "ef exit $?"; cat ${O}_ef_c4.jsonl; tail -3 ${O}_ef_c4.err
timeout 600 python scripts/probe_ef.py --config c4 --n 4000000 --latent 24 --efs 64,128,256 > ${O}_ef_c4_l24.jsonl 2> ${O}_ef_c4_l24.err; cat ${O}_ef_c4_l24.jsonl
EF=$(python -c "import json;print([json.loads(l) for l in open('${O}_ef_c4.jsonl')][-1]['chosen_ef'] or 512)")
echo "chosen ef for c4: $EF"
run() { name=$1; shift; timeout 2400 python bench.py "$@" > ${O}_$name.json 2> ${O}_$name.err; echo "$name exit $?"; tail -c 2600 ${O}_$name.json; echo; tail -3 ${O}_$name.err; }
run c3 --config c3 --steps 10
run c4 --config c4 --ef $EF --steps 10
