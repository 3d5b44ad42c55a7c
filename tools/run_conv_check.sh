#!/bin/bash
# One GPU call: conv-family kernel tests (one group per process), per-CTA phase timing, per-shape micro-benchmark.
# usage: tools/run_conv_check.sh TAG [quick]
TAG=${1:-x}
OUT=gpurun_out/r02
mkdir -p $OUT
for grp in conv2d linear qkv layernorm; do
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$grp" -p no:cacheprovider > $OUT/test_${TAG}_$grp.log 2>&1
  echo "== $grp: exit $? =="; tail -n 4 $OUT/test_${TAG}_$grp.log
done
timeout 300 python tools/conv_phases.py conv3 conv3_96 mid combos=00,01,10,11 > $OUT/phases_${TAG}.txt 2>&1
timeout 300 python tools/conv_phases.py conv3 conv3_plain conv1 combos=00 bn=64 >> $OUT/phases_${TAG}.txt 2>&1
cat $OUT/phases_${TAG}.txt
timeout 300 python tools/bench_conv.py --set step > $OUT/bench_conv_${TAG}.txt 2>&1
cat $OUT/bench_conv_${TAG}.txt
if [ "$2" != "quick" ]; then
  timeout 600 python tools/bench_conv.py --set large > $OUT/bench_conv_${TAG}_large.txt 2>&1
  cat $OUT/bench_conv_${TAG}_large.txt
fi
