#!/usr/bin/env bash
# per-workload device_ms for a few concurrency settings (tools/profile_frame.py)
cd "$(dirname "$0")/.."
for w in "$@"; do
  for c in 1 2 4 8; do
    printf "%s conc=%s " "$w" "$c"
    python tools/profile_frame.py --workload "$w" --frames 4 --concurrency "$c" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['device_ms'], d['closest_ms'], d['shadow_ms'], d['kernel_launches'])"
  done
done
