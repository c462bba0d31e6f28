#!/usr/bin/env bash
# Rebuilds one kernel file with a -D knob on the GPU box (same image: nvcc is there) and times the stages.
#   tools/tune_kernels.sh <file.cu> <MACRO> <v1> [v2 ...]      e.g.  tools/tune_kernels.sh fd_brief.cu BRIEF_UNROLL 2 4 8
set -u
cd "$(dirname "$0")/.."
f=$1; m=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  touch feature_detector_b200/csrc/$f feature_detector_b200/csrc/fd_kernels.cuh
  make -s -j16 -C feature_detector_b200/csrc EXTRA=-D$m=$v > gpurun_out/tune_build.log 2>&1 || { echo "build failed for $m=$v"; tail -5 gpurun_out/tune_build.log; continue; }
  echo "== $m=$v"
  python tools/stage_times.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
print({k:(v['candidates_ms'],v['select_ms'],v['brief_ms'],v['step_ms']) for k,v in d.items() if 'step_ms' in v}, d['one_frame_us'])"
done
touch feature_detector_b200/csrc/$f feature_detector_b200/csrc/fd_kernels.cuh
make -s -j16 -C feature_detector_b200/csrc > /dev/null 2>&1
