#!/bin/bash
# Launch list of ONE configs[3] ResNet training step + ncu --set full of its three HMMA kernels at layer1 size.
# usage (on the GPU box): bash profiles/capture_r01_resnet.sh <tag>
TAG=${1:-x}
python profiles/run_resnet_step.py > gpurun_out/rn_plain_$TAG.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/rn_launches_$TAG.csv python profiles/run_resnet_step.py > gpurun_out/rn_ncu_list_$TAG.log 2>&1
cap() {  # name regex skip count
  ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/rn_prof_$1_$TAG python profiles/run_resnet_step.py > gpurun_out/rn_ncu_$1_$TAG.log 2>&1
  tail -1 gpurun_out/rn_ncu_$1_$TAG.log
}
cap mma_fwd 'rn_conv3_mma_fwd_kernel' 0 1
cap wgrad_mma 'rn_wgrad_mma16_kernel<.int.64, .int.1, .int.14' 0 1
cap k8_dgrad 'rn_conv3_k8_mma_kernel<.int.8' 0 1
cap stem_fwd 'rn_stem_mma_fwd_kernel' 0 1
cap stem_wgrad 'rn_stem_mma_wgrad_kernel' 0 1
cap bn_bwd_apply 'rn_bn_bwd_apply_kernel<.bool.0' 6 1     # the stem's (last launch of the backward)
