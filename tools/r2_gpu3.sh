#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=$PWD/course-assignment-danielhalachev_b200/csrc/variants
echo "== default build: threshold decay / start / cap"
timeout 1200 python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,synthetic_10M --tails 16:8,16:8:1024:32768:2,16:8:1024:32768:1,16:8:256:32768:2,16:8:256:32768:1,32:8:256:32768:1,16:12:256:32768:1,16:16:256:32768:2 --shards 1,4,8 --json $O/r2m_decay.json > $O/r2m_decay.txt 2>&1; grep -v "^\[bench\]" $O/r2m_decay.txt | grep -v literal | cut -c1-215
echo "== k_coop with child prefetch"
CRT_CORE_LIB=$V/libcrtb200_pf.so timeout 900 python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,synthetic_10M --tails 16:8,16:8:256:32768:1,16:16:256:32768:2 --shards 1,4,8 --json $O/r2m_pf.json > $O/r2m_pf.txt 2>&1; grep -v "^\[bench\]" $O/r2m_pf.txt | grep -v literal | cut -c1-215
