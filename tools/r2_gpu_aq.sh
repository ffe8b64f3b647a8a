#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=r2aq
E2E_PER_SET=1,2,3 E2E_SETS=2,3,4 python tools/e2e_time.py hw14_dragon_class > $O/${T}_e2e.txt 2>&1
E2E_PER_SET=1,2 E2E_SETS=1,2,4 python tools/e2e_time.py hw11_room >> $O/${T}_e2e.txt 2>&1
E2E_PER_SET=1,2 E2E_SETS=1,2,4 python tools/e2e_time.py hw07_scene0b >> $O/${T}_e2e.txt 2>&1
grep -v "^\[bench" $O/${T}_e2e.txt
python tools/r2_measure.py --workloads hw14_dragon_class,synthetic_10M,hw11_room --tails 16:8:512::512,16:8:512::256,16:8:512::128,16:8:1024::256,16:8:512::64 --shards 1,8 --frames 9 2>&1 | grep -v "^\[bench\]\|literal" | tee $O/${T}_policy.txt
