#!/usr/bin/env bash
# time the default build, the 4-wide walk (CRT_LAYOUT=wide) and every tuning build under csrc/variants/
# (tools only; the product loads csrc/libcrtb200.so)
cd "$(dirname "$0")/.."
t() { python tools/profile_frame.py --workload "$1" --frames 4 --concurrency 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-20s %8.3f %8.3f %8.3f' % ('$1', d['device_ms'], d['closest_ms'], d['shadow_ms']))"; }
echo "== default"; for w in "$@"; do t $w; done
if [ -n "$CRT_AB_WIDE" ]; then echo "== default, CRT_LAYOUT=wide"; for w in "$@"; do CRT_LAYOUT=wide t $w; done; fi
for so in course-assignment-danielhalachev_b200/csrc/variants/*.so; do
  [ -e "$so" ] || continue
  echo "== $(basename $so)"
  for w in "$@"; do CRT_CORE_LIB=$PWD/$so t $w; done
done
