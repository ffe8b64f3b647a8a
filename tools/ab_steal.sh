#!/usr/bin/env bash
# A/B of range stealing (CRT_STEAL) for whole frames and for 1/N tile shards -- tools only
cd "$(dirname "$0")/.."
t() { python tools/profile_frame.py --workload "$1" --frames 5 --concurrency 1 --shards "$2" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-20s shards %s  %8.3f %8.3f %8.3f' % ('$1', '$2', d['device_ms'], d['closest_ms'], d['shadow_ms']))"; }
for st in 1 0; do echo "== CRT_STEAL=$st"; for w in "$@"; do for s in 1 8; do CRT_STEAL=$st t $w $s; done; done; done
