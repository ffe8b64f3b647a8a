#!/usr/bin/env bash
# multi-GPU measurements on one box (run under `gpurun --gpus 8`): the driver's own launch line for N = 1, 2, 4, 8
# (tiles = default), the C-ABI multi-GPU context (tests + timing), and the config-5 animation at N = 8.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; export TAG=${1:-r2s}
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/${TAG}_gpus.txt
python bench.py --steps 10 --warmup 3 --no-config5 > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err; tail -c 400 $O/${TAG}_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + N)) bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_n${N}.json 2> $O/${TAG}_n${N}.err
  tail -c 300 $O/${TAG}_n${N}.err | tail -2
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --workload synthetic_10M --animation 60 --steps 3 --warmup 1 > $O/${TAG}_n8_anim60.json 2> $O/${TAG}_n8_anim60.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi or binding" > $O/${TAG}_pytest_multi.log 2>&1; tail -3 $O/${TAG}_pytest_multi.log
timeout 600 python tools/multi_ctx_time.py > $O/${TAG}_multi_ctx.txt 2>&1; cat $O/${TAG}_multi_ctx.txt
python - <<'PY'
import json, glob
import os
for f in sorted(glob.glob("gpurun_out/" + os.environ.get("TAG", "r2s") + "_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), "value %.1f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.1f" % d.get("e2e", {}).get("value", 0), d.get("tiles_check"), (d.get("frames_mode") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
