#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=r2ae
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 800 > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,hw11_room_128,synthetic_10M --tails 16 --shards 1,8 --frames 9 2>&1 | grep -v "^\[bench\]\|literal" | tee $O/${T}_measure.txt
python tools/chunk_times.py hw14_dragon_class > $O/${T}_chunks.txt 2>&1
CRT_CHUNK_STAGGER=0 python tools/chunk_times.py hw14_dragon_class >> $O/${T}_chunks.txt 2>&1
CRT_HOST_CHUNKS_PER_SET=3 E2E_SETS=3 python tools/chunk_times.py hw14_dragon_class >> $O/${T}_chunks.txt 2>&1
E2E_STAGGERS=0,0.15 E2E_PER_SET=2,3 E2E_SETS=3,4 python tools/e2e_time.py hw14_dragon_class > $O/${T}_e2e.txt 2>&1
grep -v "^\[bench" $O/${T}_chunks.txt | tail -40; grep -v "^\[bench" $O/${T}_e2e.txt
