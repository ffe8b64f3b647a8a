#!/usr/bin/env bash
# multi-GPU bench lines at N ranks (run under `gpurun --gpus N`): frames (weak), tiles (strong), config-5 animation
cd "$(dirname "$0")/.."
N=$1; TAG=${2:-scale}; O=gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > $O/${TAG}_n${N}_$2.json 2> $O/${TAG}_n${N}_$2.err; tail -c 300 $O/${TAG}_n${N}_$2.err | tail -2; }
run 29541 frames --steps 10 --warmup 3
run 29542 tiles --steps 10 --warmup 3 --parallelism tiles
run 29543 anim60_10M --workload synthetic_10M --animation 60 --steps 3 --warmup 1
