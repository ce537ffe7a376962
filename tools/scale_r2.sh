#!/bin/bash
# bench.py at N GPUs under torchrun, as the driver launches it; usage: tools/scale_r2.sh N
cd "$(dirname "$0")/.."
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/r2i_scale_n$N.json 2> gpurun_out/r2i_scale_n$N.err
tail -c 400 gpurun_out/r2i_scale_n$N.json
