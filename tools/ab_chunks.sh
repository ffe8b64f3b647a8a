#!/usr/bin/env bash
# device-only frames split into k chunks on k streams (CRT_DEVICE_CHUNKS): do the tails of one chunk's persistent kernels get filled by the next chunk's?
cd "$(dirname "$0")/.."
t() { python tools/profile_frame.py --workload "$1" --frames 6 --concurrency 8 --shards "$2" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-20s shards %s  %8.3f ms  launches %d' % ('$1', '$2', d['device_ms'], d['kernel_launches']))"; }
for k in 1 2 3 4 6 8; do echo "== CRT_DEVICE_CHUNKS=$k"; for w in "$@"; do for s in 1 8; do CRT_DEVICE_CHUNKS=$k t $w $s; done; done; done
