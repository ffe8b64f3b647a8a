cd /root/repo
echo "== default (m0 r8)"; python tools/profile_frame.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['device_ms'], d['closest_ms'], d['shadow_ms'])"
python tools/profile_frame.py --workload hw11_room 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['device_ms'], d['closest_ms'], d['shadow_ms'])"
for v in m0_r4 m0_r16 m1_r4 m1_r8 m1_r16; do
  echo "== $v"
  for w in hw14_dragon_class hw11_room; do
    CRT_CORE_LIB=/root/repo/course-assignment-danielhalachev_b200/csrc/variants/libcrtb200_$v.so python tools/profile_frame.py --workload $w 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['device_ms'], d['closest_ms'], d['shadow_ms'])"
  done
done
