#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_run45
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_c2_n2.json 2> ${O}_c2_n2.err; echo "n2 exit $?"; tail -c 1500 ${O}_c2_n2.json; tail -3 ${O}_c2_n2.err
