#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=$PWD/course-assignment-danielhalachev_b200/csrc/variants
for w in hw11_room; do
echo "== $w (phase clocks build, hand-off off, then default)"
CRT_CORE_LIB=$V/libcrtb200_pc.so CRT_WARP_DUMP=1 CRT_TAIL_ITERS=-1 python tools/profile_frame.py --workload $w --frames 2 --concurrency 1 2>&1 | tail -60
CRT_CORE_LIB=$V/libcrtb200_pc.so CRT_WARP_DUMP=1 python tools/profile_frame.py --workload $w --frames 2 --concurrency 1 2>&1 | tail -60
done > $O/r2q_phase.txt 2>&1
cut -c1-330 $O/r2q_phase.txt
