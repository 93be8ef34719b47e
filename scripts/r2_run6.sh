#!/bin/bash
# round 2, GPU run 6 (8 GPUs): c2 weak / strong with the library's fused gather, c3 and c4 at their stated sizes
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out/r2_run6_n$N
tr() { name=$1; shift; timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > ${O}_$name.json 2> ${O}_$name.err; echo "$name exit $?"; grep "^{" ${O}_$name.json | tail -1 | cut -c1-700; grep -v "^$\|OMP_NUM\|\*\*\*\*" ${O}_$name.err | tail -4; }
tr c2_p2p --config c2 --steps 20 --warmup 3 --no-cpu-baseline
tr c2_nccl --config c2 --steps 20 --warmup 3 --no-cpu-baseline --gather nccl
tr c2_strong --config c2 --steps 20 --warmup 3 --no-cpu-baseline --strong --nq 8192
tr c3 --config c3 --steps 10 --warmup 3 --no-cpu-baseline
tr c4 --config c4 --steps 10 --warmup 3 --no-cpu-baseline
