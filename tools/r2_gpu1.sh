#!/usr/bin/env bash
# round-2 GPU call 1: parity suite (incl. the full-size reference-binary tests) + the traversal / hand-off A/B matrix
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 1400 -x --deselect tests/test_gpu_fullsize.py::test_10m_frame_identical_to_reference_binary > $O/r2a_pytest.log 2>&1; tail -15 $O/r2a_pytest.log
timeout 900 python tools/r2_measure.py --workloads hw14_dragon_class,hw11_room,synthetic_10M --tails 16:4,32:1,8:8,4:16,0:0 --shards 1,8 --json $O/r2a_matrix.json > $O/r2a_matrix.txt 2>&1; cat $O/r2a_matrix.txt | grep -v "^\[bench\]"
timeout 1400 python -m pytest tests/test_gpu_fullsize.py::test_10m_frame_identical_to_reference_binary -q --timeout 1300 > $O/r2a_pytest_10m.log 2>&1; tail -5 $O/r2a_pytest_10m.log
