#!/bin/bash
# One kernel-iteration round on the GPU box: bit-identity + timing sweep, then one ncu full capture of the streaming kernel.
# usage: tools/gpu_iter.sh <tag> [sweep args...]   (env: NCU=0/1, PROF_CONFIG, KERNEL)
tag=$1; shift
mkdir -p gpurun_out
SWEEP_ITERS=${SWEEP_ITERS:-10} timeout 600 python tools/sweep_configs.py "$@" > gpurun_out/${tag}_sweep.log 2>&1
cat gpurun_out/${tag}_sweep.log
if [ "${NCU:-1}" = "1" ]; then
  cfg=${PROF_CONFIG:-10}; kern=${KERNEL:-stream4_kernel}
  PROF_CONFIG=$cfg PROF_LAUNCHES=3 python tools/prof_fused.py > gpurun_out/plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -o gpurun_out/${tag} -f env PROF_CONFIG=$cfg PROF_LAUNCHES=3 python tools/prof_fused.py > gpurun_out/${tag}_ncu.log 2>&1
  tail -2 gpurun_out/${tag}_ncu.log
fi
