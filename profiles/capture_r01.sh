#!/bin/bash
# ncu --set full captures of single launches of the configs[1] training step (profiles/run_step.py), block-1 shapes,
# plus the launch list of the whole step.  usage (on the GPU box): bash profiles/capture_r01.sh <tag>
TAG=${1:-x}
python profiles/run_step.py cfg2 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$TAG.csv python profiles/run_step.py cfg2 > gpurun_out/ncu_list_$TAG.log 2>&1
cap() {  # name regex skip count
  ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/prof_$1_$TAG python profiles/run_step.py cfg2 > gpurun_out/ncu_$1_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_$1_$TAG.log
}
cap wgrad 'conv_wgrad_kernel<.int.0, .int.1' 112 2          # last two of 116: conv2 + conv1 weight gradient of block 1
cap brick_fprop 'conv3_brick_kernel<.int.1' 3 1            # block 1, layer 4
cap brick_dgrad 'conv3_brick_kernel<.int.0' 55 1           # block 1
cap stem 'stem_brick_kernel' 0 1
cap rows_dgrad 'conv_rows_kernel<.int.0, .int.0, .int.2, .bool.1, .int.1' 30 1
cap bnbwd 'bn_bwd_apply_kernel<.int.1' 55 1
