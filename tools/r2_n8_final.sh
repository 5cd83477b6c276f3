#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tools/h2d_probe.py > gpurun_out/r2_h2d_probe_n$N.md 2>&1; tail -8 gpurun_out/r2_h2d_probe_n$N.md
bash tools/r2_nN.sh $N p2p
