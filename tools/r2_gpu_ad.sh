#!/usr/bin/env bash
# scratch: early k_coop pass A/B + host-frame chunk timeline
cd "$(dirname "$0")/.."
O=gpurun_out; T=r2ad
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 800 > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
for e in 0 2 3 6; do
  echo "== CRT_COOP_EARLY=$e" | tee -a $O/${T}_early.txt
  CRT_COOP_EARLY=$e python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,synthetic_10M --tails 16 --shards 1,8 --frames 9 2>&1 | grep -v "^\[bench\]\|literal" | tee -a $O/${T}_early.txt
done
python tools/chunk_times.py hw14_dragon_class > $O/${T}_chunks.txt 2>&1
CRT_HOST_CHUNKS_PER_SET=3 E2E_SETS=3 python tools/chunk_times.py hw14_dragon_class >> $O/${T}_chunks.txt 2>&1
tail -30 $O/${T}_chunks.txt
