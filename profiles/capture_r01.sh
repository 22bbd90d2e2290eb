#!/bin/bash
# ncu --set full captures of single launches of the configs[1] training step (profiles/run_step.py), block-1 shapes.
# usage (on the GPU box): bash profiles/capture_r01.sh <tag>
TAG=${1:-x}
cap() {  # name regex skip count
  ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/prof_$1_$TAG python profiles/run_step.py cfg2 > gpurun_out/ncu_$1_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_$1_$TAG.log
}
cap brick_dgrad 'conv3_brick_kernel<.int.0' 55 1
cap wgrad 'conv_wgrad_kernel<.int.0, .int.1' 112 2
cap rows_dgrad 'conv_rows_kernel<.int.0, .int.0, .int.2, .bool.1, .int.1' 30 1
cap rows_fprop 'conv_rows_kernel<.int.0, .int.1, .int.1, .bool.0, .int.1' 3 1
cap bnbwd 'bn_bwd_apply_kernel<.int.1' 55 1
cap stem 'conv_rows_kernel<.int.1' 0 1
