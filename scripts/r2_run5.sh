#!/bin/bash
# round 2, GPU run 5 (1 GPU): ncu captures (launch list of the bench's timed region, full sets of the search and GEMM
# kernels), BM25 timing, harder data rows for c2
mkdir -p gpurun_out
O=gpurun_out/r2_run5
NCU="ncu --clock-control none"
timeout 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum -c 200 --csv --log-file ${O}_launches_c2.csv python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > ${O}_launches_c2.log 2>&1; echo "launch list exit $?"
timeout 900 $NCU --profile-from-start off --set full --import-source on -k regex:hnsw_search_kernel --launch-skip 3 --launch-count 1 -f -o ${O}_prof_search_c2 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > ${O}_prof_search_c2.log 2>&1; echo "search ncu exit $?"
timeout 900 $NCU --set full --import-source on -k regex:gemm_tc_kernel --launch-skip 5 --launch-count 1 -f -o ${O}_prof_gemm python scripts/probe_gemm.py --n 1000000 --skip-exact > ${O}_prof_gemm.log 2>&1; echo "gemm ncu exit $?"
timeout 900 $NCU --set full --import-source on -k regex:bm25_hash_kernel --launch-skip 2 --launch-count 1 -f -o ${O}_prof_bm25 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > ${O}_prof_bm25.log 2>&1; echo "bm25 ncu exit $?"
for L in 32 64 768; do timeout 600 python scripts/probe_ef.py --config c2 --latent $L --efs 64,128,256,512,1024 >> ${O}_ef_c2_latents.jsonl 2>> ${O}_ef_c2_latents.err; done; cat ${O}_ef_c2_latents.jsonl
ls -la gpurun_out | grep r2_run5
