#!/bin/bash
# round 2, GPU run 4 (N GPUs): the library's fused peer-store gather against NCCL, weak and strong scaling on c2
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r2_run4_n$N
nvidia-smi topo -m > ${O}_topo.txt 2>&1
tr() { name=$1; shift; NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > ${O}_$name.json 2> ${O}_$name.err; echo "$name exit $?"; tail -c 1500 ${O}_$name.json; echo; grep -v "^$" ${O}_$name.err | tail -4; }
tr p2p --config c2 --steps 20 --warmup 3 --no-cpu-baseline
tr nccl --config c2 --steps 20 --warmup 3 --no-cpu-baseline --gather nccl
tr strong --config c2 --steps 20 --warmup 3 --no-cpu-baseline --strong --nq 8192
