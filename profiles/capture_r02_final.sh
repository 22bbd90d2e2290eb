#!/bin/bash
# Round-2 final evidence: default bench line, ncu launch list of ONE configs[1] training step, ncu --set full of the
# TMA-fed weight gradients (block-1 3x3x3 + 1x1x1, stem).  usage (on the GPU box): bash profiles/capture_r02_final.sh <tag>
TAG=${1:-x}
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
python profiles/run_step.py cfg2 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$TAG.csv python profiles/run_step.py cfg2 > gpurun_out/ncu_list_$TAG.log 2>&1
cap() {  # name regex skip count
  ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/prof_$1_$TAG python profiles/run_step.py cfg2 > gpurun_out/ncu_$1_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_$1_$TAG.log
  python profiles/ncu_summary.py gpurun_out/prof_$1_$TAG.ncu-rep > gpurun_out/ncu_summary_$1_$TAG.txt 2>&1
  python profiles/ncu_hot.py gpurun_out/prof_$1_$TAG.ncu-rep 30 > gpurun_out/ncu_hot_$1_$TAG.txt 2>&1
  rm -f gpurun_out/prof_$1_$TAG.ncu-rep
}
cap wgrad 'conv_wgrad_kernel<.int.0, .int.1' 112 2
cap stem_wgrad 'conv_wgrad_kernel<.int.1' 0 1
python profiles/timeline.py > gpurun_out/timeline_$TAG.txt 2>&1
