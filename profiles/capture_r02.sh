#!/bin/bash
# Round-2 evidence: ncu launch list of ONE configs[1] training step + ncu --set full of single block-1 launches.
# usage (on the GPU box): bash profiles/capture_r02.sh <tag>
TAG=${1:-x}
python profiles/run_step.py cfg2 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$TAG.csv python profiles/run_step.py cfg2 > gpurun_out/ncu_list_$TAG.log 2>&1
cap() {  # name regex skip count
  ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/prof_$1_$TAG python profiles/run_step.py cfg2 > gpurun_out/ncu_$1_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_$1_$TAG.log
  # digest on the box (the raw reports are ~13 MB each; gpurun_out/ is capped at 64 MiB) and drop the report
  python profiles/ncu_summary.py gpurun_out/prof_$1_$TAG.ncu-rep > gpurun_out/ncu_summary_$1_$TAG.txt 2>&1
  python profiles/ncu_hot.py gpurun_out/prof_$1_$TAG.ncu-rep 30 > gpurun_out/ncu_hot_$1_$TAG.txt 2>&1
  rm -f gpurun_out/prof_$1_$TAG.ncu-rep
}
cap wgrad 'conv_wgrad_kernel<.int.0, .int.1' 112 2          # last two of 116: conv2 + conv1 weight gradient of block 1 (TMA-fed)
cap rows_dgrad_acc 'conv_rows_kernel<.int.0, .int.0, .int.3' 55 1   # 1x1x1 data gradient with the accumulate epilogue, block 1
cap brick_fprop 'conv3_brick_kernel<.int.1' 3 1             # block 1, layer 4
cap finalize 'grad_finalize_kernel' 55 1                     # block 1 slice
cap stem_wgrad 'conv_wgrad_kernel<.int.1' 0 1
