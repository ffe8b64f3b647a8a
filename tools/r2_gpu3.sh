#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=$PWD/course-assignment-danielhalachev_b200/csrc/variants
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 800 > $O/r2g_pytest.log 2>&1; tail -5 $O/r2g_pytest.log
timeout 1200 python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,hw11_room_128,synthetic_10M --tails 16:1,16:2,16:4,16:8,64:2,64:4,4:2,16:2:256,16:4:4096,-1 --shards 1,8 --json $O/r2g_matrix.json > $O/r2g_matrix.txt 2>&1; grep -v "^\[bench\]" $O/r2g_matrix.txt | grep -v literal
echo "== g8"
CRT_CORE_LIB=$V/libcrtb200_g8.so timeout 600 python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,synthetic_10M --tails 16:2,16:8 --shards 1,8 > $O/r2g_matrix_g8.txt 2>&1; grep -v "^\[bench\]" $O/r2g_matrix_g8.txt | grep -v literal
