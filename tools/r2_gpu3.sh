#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=$PWD/course-assignment-danielhalachev_b200/csrc/variants
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 800 -x > $O/r2h_pytest.log 2>&1; tail -3 $O/r2h_pytest.log
echo "== k_wave"
timeout 1200 python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,hw11_room_128,synthetic_10M --tails 16:8,16:64:16:100000000,16:64:1024:100000000,64:64:64:100000000,16:8:1024:65536,-1 --shards 1,8 --json $O/r2h_wave.json > $O/r2h_wave.txt 2>&1; grep -v "^\[bench\]" $O/r2h_wave.txt | grep -v literal | cut -c1-215
echo "== k_coop"
CRT_CORE_LIB=$V/libcrtb200_coop.so timeout 1200 python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,hw11_room_128,synthetic_10M --tails 16:8,16:64:16:100000000,-1 --shards 1,8 --json $O/r2h_coop.json > $O/r2h_coop.txt 2>&1; grep -v "^\[bench\]" $O/r2h_coop.txt | grep -v literal | cut -c1-215
