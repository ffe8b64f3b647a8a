#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=$PWD/course-assignment-danielhalachev_b200/csrc/variants
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 800 -k "multi or context_error or refuses or assemble or small_chunks" > $O/r2f_pytest.log 2>&1; tail -15 $O/r2f_pytest.log
for v in g32s g8s; do echo "== $v"; CRT_CORE_LIB=$V/libcrtb200_$v.so timeout 600 python tools/r2_coopstats.py hw14_dragon_class,synthetic_10M 8,64 2>&1 | grep -v "^\[bench\]" | awk '/coop stats/{last=$0; next} {if(last!=""){print last; last=""} print}'; done > $O/r2f_coopstats.txt 2>&1; cat $O/r2f_coopstats.txt
for v in g32 g16; do
echo "== variant $v"
CRT_CORE_LIB=$V/libcrtb200_$v.so timeout 600 python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,synthetic_10M --tails 8,64 --shards 1,8 > $O/r2f_matrix_$v.txt 2>&1; grep -v "^\[bench\]" $O/r2f_matrix_$v.txt | grep -v literal
done
