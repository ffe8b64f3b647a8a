#!/usr/bin/env bash
# round-2 GPU call 2: hand-off variants (iteration threshold), main / coop kernel split, shards 1 / 2 / 4 / 8
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 > $O/r2b_pytest.log 2>&1; tail -5 $O/r2b_pytest.log
timeout 1200 python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,hw11_room_128,synthetic_10M --tails 8,16,32,64,128,-1 --shards 1,2,4,8 --json $O/r2b_matrix.json > $O/r2b_matrix.txt 2>&1; grep -v "^\[bench\]" $O/r2b_matrix.txt
