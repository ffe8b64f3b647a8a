#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q --timeout 1500 -x > $O/r2t_pytest.log 2>&1; tail -3 $O/r2t_pytest.log
timeout 900 python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,hw11_room_128,synthetic_10M,hw07_scene0b,hw12_textures --tails 16:8,-1 --shards 1,8 --json $O/r2t_matrix.json > $O/r2t_matrix.txt 2>&1; grep -v "^\[bench\]" $O/r2t_matrix.txt | cut -c1-215
