#!/usr/bin/env bash
# A/B of the L2 access-policy window modes (CRT_L2_PERSIST, crtb200_core.cu) -- tools only
cd "$(dirname "$0")/.."
t() { python tools/profile_frame.py --workload "$1" --frames 5 --concurrency 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-20s %8.3f %8.3f %8.3f' % ('$1', d['device_ms'], d['closest_ms'], d['shadow_ms']))"; }
for rep in 1 2; do for m in 0 1 2 3; do echo "== CRT_L2_PERSIST=$m"; for w in "$@"; do CRT_L2_PERSIST=$m t $w; done; done; done
