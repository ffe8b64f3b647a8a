#!/usr/bin/env bash
# Round-end measurement set on one B200 (run under gpurun): parity tests, one bench line per workload, the ncu launch
# list of the default bench command and one `--set full` capture of the traversal kernels.  Outputs: gpurun_out/$TAG_*
cd "$(dirname "$0")/.."
TAG=${1:-final}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; tail -2 $O/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_hw14_dragon_class.json 2> $O/${TAG}_bench_hw14_dragon_class.err
for w in hw07_scene0b hw11_room hw11_room_128 hw12_textures; do
  python bench.py --workload $w --steps 10 --warmup 3 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
done
python bench.py --workload synthetic_10M --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_synthetic_10M.json 2> $O/${TAG}_bench_synthetic_10M.err
python bench.py --workload synthetic_10M --animation 60 --steps 3 --warmup 1 > $O/${TAG}_anim60_synthetic_10M.json 2> $O/${TAG}_anim60.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err
# ncu: launch list of the bench command, then the full capture (each after its plain run exited 0)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
for w in hw14_dragon_class synthetic_10M; do
  python tools/profile_frame.py --workload $w --frames 2 --concurrency 1 > $O/${TAG}_plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"k_closest|k_shadow" -s 2 -c 2 -f -o $O/${TAG}_prof_$w \
      python tools/profile_frame.py --workload $w --frames 2 --concurrency 1 > $O/${TAG}_ncu_$w.log 2>&1
done
ls -la $O | grep ${TAG}_ | awk '{print $5, $9}'
