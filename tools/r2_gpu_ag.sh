#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=r2ag
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 1400 > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
python tools/r2_measure.py --workloads hw14_dragon_class,synthetic_10M,hw11_room --tails 16:8:1024,16:8:512,16:8:256,16:16:1024,16:16:512,8:8:1024,32:8:1024 --shards 1,8 --frames 9 2>&1 | grep -v "^\[bench\]\|literal" | tee $O/${T}_policy.txt
