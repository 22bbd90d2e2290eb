#!/bin/bash
# ncu --set full captures of LATE-BLOCK launches (block 4: M = 512 rows, 4 row tiles) of the configs[1] training step.
TAG=${1:-x}
python profiles/run_step.py cfg2 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
cap() {  # name regex skip count
  ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/prof_$1_$TAG python profiles/run_step.py cfg2 > gpurun_out/ncu_$1_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_$1_$TAG.log
}
cap rows_fprop_b4 'conv_rows_kernel<.int.0, .int.1, .int.1, .bool.0, .int.2' 39 1
cap brick_fprop_b4 'conv3_brick_kernel<.int.1' 57 1
cap rows_dgrad_b4 'conv_rows_kernel<.int.0, .int.0, .int.2, .bool.1, .int.2' 0 1
cap brick_dgrad_b4 'conv3_brick_kernel<.int.0' 0 1
