#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "in_place or multi" > $O/r2n_pytest.log 2>&1; tail -3 $O/r2n_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2n_n2.json 2> $O/r2n_n2.err; tail -5 $O/r2n_n2.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2n_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","ms_per_step","n_gpus","scaling")}, d.get("tiles_check"), d.get("tiles_other_transport"), d.get("frames_mode",{}).get("value"), d.get("e2e"))
print(d["config"]["parallelism"])
PY
