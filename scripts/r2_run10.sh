#!/bin/bash
# round 2, GPU run 10 (N GPUs): the binary config's 1 -> 8 sweep, N = 2 and 4 (N = 1 and 8 are r2e / r2d)
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r2_run10_n$N
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config c4 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_c4.json 2> ${O}_c4.err; echo "c4 exit $?"; grep "^{" ${O}_c4.json | tail -1 | cut -c1-400; grep -v "^$\|OMP_NUM\|\*\*\*\*" ${O}_c4.err | tail -4
