#!/usr/bin/env bash
# default build vs every build under csrc/variants/, whole frame and 1 of 8 tile shards -- tools only
cd "$(dirname "$0")/.."
t() { python tools/profile_frame.py --workload "$1" --frames 5 --concurrency 1 --shards "$2" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-20s shards %s  %8.3f %8.3f %8.3f' % ('$1', '$2', d['device_ms'], d['closest_ms'], d['shadow_ms']))"; }
echo "== default"; for w in "$@"; do for s in 1 8; do t $w $s; done; done
for so in course-assignment-danielhalachev_b200/csrc/variants/*.so; do [ -e "$so" ] || continue; echo "== $(basename $so)"; for w in "$@"; do for s in 1 8; do CRT_CORE_LIB=$PWD/$so t $w $s; done; done; done
