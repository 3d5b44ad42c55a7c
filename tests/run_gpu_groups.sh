#!/bin/bash
# Runs the GPU kernel tests one group per process (a device-side trap poisons the CUDA context of its process only).
mkdir -p gpurun_out
for grp in conv2d linear qkv groupnorm layernorm attention elementwise cfg_ddim; do
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" -p no:cacheprovider > gpurun_out/test_$grp.log 2>&1
  echo "== $grp: exit $? =="; tail -n 25 gpurun_out/test_$grp.log
done
