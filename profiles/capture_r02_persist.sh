#!/bin/bash
# ncu --set full of the block-1 persistent 1x1x1 forward GEMM (last layer of block 1, Cin = 224) inside one configs[1] training step
TAG=${1:-x}
python profiles/run_step.py cfg2 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled \
    -k "regex:conv1_persist_kernel" -s 5 -c 1 -f -o gpurun_out/prof_persist_$TAG python profiles/run_step.py cfg2 > gpurun_out/ncu_persist_$TAG.log 2>&1
tail -1 gpurun_out/ncu_persist_$TAG.log
python profiles/ncu_summary.py gpurun_out/prof_persist_$TAG.ncu-rep > gpurun_out/ncu_summary_persist_$TAG.txt 2>&1
python profiles/ncu_hot.py gpurun_out/prof_persist_$TAG.ncu-rep 40 > gpurun_out/ncu_hot_persist_$TAG.txt 2>&1
rm -f gpurun_out/prof_persist_$TAG.ncu-rep
