#!/usr/bin/env bash
# Round-end measurement set on one B200 (run under gpurun): parity tests, one bench line per workload, the reference arm,
# the ncu launch list of the default bench command and one `--set full` capture of the traversal kernels.
# Outputs: gpurun_out/$TAG_*
cd "$(dirname "$0")/.."
TAG=${1:-final}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $O/${TAG}_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 1700 python -m pytest tests -m gpu -q --timeout 1500 > $O/${TAG}_pytest.log 2>&1; tail -3 $O/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_hw14_dragon_class.json 2> $O/${TAG}_bench_hw14_dragon_class.err; tail -c 600 $O/${TAG}_bench_hw14_dragon_class.err
for w in hw07_scene0b hw11_room hw11_room_128 hw12_textures; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-config5 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
done
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err
# ncu: launch list of the bench command, then the full captures (each after its plain run exited 0); a frame has six
# traversal launches: k_closest + two k_coop passes, k_shadow + two k_coop passes
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config5 > $O/${TAG}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config5 > $O/${TAG}_ncu_bench.log 2>&1
for w in hw14_dragon_class synthetic_10M; do
  python tools/profile_frame.py --workload $w --frames 2 --concurrency 1 > $O/${TAG}_plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"k_closest|k_shadow|k_coop" -s 6 -c 6 -f -o $O/${TAG}_prof_$w \
      python tools/profile_frame.py --workload $w --frames 2 --concurrency 1 > $O/${TAG}_ncu_$w.log 2>&1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[-1][:-5], "value %.1f" % d.get("value", 0), "ms %.3f" % d.get("ms_per_step", 0), "e2e %.1f" % (d.get("e2e") or {}).get("value", 0),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "diff px", (d.get("cpu_baseline") or {}).get("pixels_differing_from_ours"),
              "roofline frac", (d.get("roofline") or {}).get("frac"), "config5", ((d.get("config5") or {}).get("static_frame") or {}).get("value"), ((d.get("config5") or {}).get("orbit_60") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
ls -la $O | grep ${TAG}_ | awk '{print $5, $9}'
