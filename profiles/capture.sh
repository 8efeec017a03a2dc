#!/bin/bash
# Run on the GPU box:  gpurun -- 'bash profiles/capture.sh <tag> [kernels...]'
# 1) plain run of the bench command, 2) ncu launch list of the same command (last step only),
# 3) ncu --set full of the named kernels (launches of the 2nd step).
set -u
TAG=${1:-r1}; shift
KERNELS=${@:-onesweep_pass_kernel emit_pairs_kernel search_count_kernel pack_keys_kernel}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-frame --no-secondary"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 2 \
      -f -o gpurun_out/prof_${TAG}_$K $CMD > gpurun_out/ncu_${TAG}_$K.log 2>&1
done
ls gpurun_out | grep $TAG
