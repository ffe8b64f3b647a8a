#!/usr/bin/env bash
# second pass for long shadow walks (CRT_LONG_BUDGET) on 1/N tile shards of several scenes -- tools only
cd "$(dirname "$0")/.."
t() { python tools/profile_frame.py --workload "$1" --frames 5 --concurrency 1 --shards "$2" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-20s shards %s  %8.3f %8.3f %8.3f' % ('$1', '$2', d['device_ms'], d['closest_ms'], d['shadow_ms']))"; }
for b in 0 64 96; do echo "== CRT_LONG_BUDGET=$b"; for w in hw11_room synthetic_10M hw07_scene0b; do for s in 4 8; do CRT_LONG_BUDGET=$b t $w $s; done; done; CRT_LONG_BUDGET=$b t hw14_dragon_class 2; CRT_LONG_BUDGET=$b t hw14_dragon_class 4; done
