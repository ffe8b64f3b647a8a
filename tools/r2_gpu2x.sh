#!/usr/bin/env bash
# 2-GPU call: e2e chunking sweep (GPU 0), bench at N = 2 through the driver's launch line, multi-GPU context tests + timing
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python tools/e2e_time.py hw14_dragon_class > $O/r2k_e2e_sweep.txt 2>&1; grep -v "^\[bench\]" $O/r2k_e2e_sweep.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2k_n2.json 2> $O/r2k_n2.err; tail -5 $O/r2k_n2.err; cut -c1-1500 $O/r2k_n2.json
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi or binding" > $O/r2k_pytest_multi.log 2>&1; tail -3 $O/r2k_pytest_multi.log
timeout 600 python tools/multi_ctx_time.py > $O/r2k_multi_ctx.txt 2>&1; grep -v "^\[bench\]" $O/r2k_multi_ctx.txt
