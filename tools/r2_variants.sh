#!/usr/bin/env bash
# A/B of library builds on one box (tools only): every build/variants/*.so named on the command line (and the in-tree
# library as "tree") runs tools/r2_measure.py in its own process.   bash tools/r2_variants.sh TAG "workloads" v1 v2 ...
cd "$(dirname "$0")/.."
TAG=$1; WL=$2; shift 2
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/${TAG}_gpu.txt
for v in tree "$@"; do
  if [ "$v" = tree ]; then unset CRT_CORE_LIB; else export CRT_CORE_LIB=$PWD/build/variants/$v.so; fi
  echo "== $v" | tee -a $O/${TAG}_variants.txt
  python tools/r2_measure.py --workloads $WL --tails ${TAILS:-16,-1} --shards ${SHARDS:-1,8} --frames 7 2>&1 | grep -v "^\[bench\]\|${SKIP:-zzzz}" | tee -a $O/${TAG}_variants.txt
done
