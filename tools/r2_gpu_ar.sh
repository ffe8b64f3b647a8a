#!/usr/bin/env bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=r2ar
E2E_PER_SET=0,1,2 E2E_SETS=2,4 python tools/e2e_time.py hw14_dragon_class > $O/${T}_e2e.txt 2>&1
E2E_PER_SET=0,2 E2E_SETS=2,4 python tools/e2e_time.py hw11_room >> $O/${T}_e2e.txt 2>&1
E2E_PER_SET=0,2 E2E_SETS=2,4 python tools/e2e_time.py hw07_scene0b >> $O/${T}_e2e.txt 2>&1
E2E_PER_SET=0 E2E_SETS=2 python tools/e2e_time.py hw12_textures >> $O/${T}_e2e.txt 2>&1
grep -v "^\[bench" $O/${T}_e2e.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 800 -k "not 10m" > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
