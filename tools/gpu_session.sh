#!/usr/bin/env bash
# One-liners for a GPU box (run under `gpurun -- 'tools/gpu_session.sh <what> [...]'`); everything lands in gpurun_out/.
# Measured costs of round 1 (warm box): tests ~75 s, bench ~25 s, bench --no-extras ~15 s, guard ~8 s, pipes ~2 s.
#   tests        python -m pytest tests -m gpu -x -q            (the round-end suite)
#   fuzz         the fuzz of the C ABI against the checker on frames from 1 x 1 pixels up alone (it is part of `tests`)
#   guard        every kernel instantiation with red zones round every buffer (tools/sanitize_paths.py)
#   bench        python bench.py  -> gpurun_out/bench_<tag>.json          (tag = $2, default "run")
#   bench-n N    torchrun bench.py --gpus N (use with gpurun --gpus N)    (tag = $3)
#   nccl         the multi-GPU tests (sharded batch + gather, row-tiled 4K frame); use with gpurun --gpus 2
#   launches     ncu launch list of the headline step (gpu__time_duration.sum, --clock-control none)
#   ncu          one --set full capture of the headline step's three kernels -> gpurun_out/prof_<tag>.ncu-rep, then read it HERE with
#                tools/ncu_summary.py / tools/ncu_regions.py / tools/ncu_hotspots.py
#   pipes        tools/microbench/pipes (issue rates of the instruction mixes the kernels are made of)
#   stages       tools/stage_times.py: FAST / selection / BRIEF and the corner detectors stage by stage, one frame and 1024 frames
#   profile-r2   the round-2 evidence set: launch list + ncu --set full of the headline step and of configs 2-4 (tools/prof_r2.py), one
#                report per kernel group so that gpurun_out/ stays under gpurun's 64 MiB; read back HERE with tools/ncu_summary.py,
#                tools/ncu_regions.py, tools/make_traffic_json.py (-> profiles/)
#   one-frame    tools/exp_single_latency.py: every piece of the one-frame-per-call path timed alone (FD_EXP_KN=9|12)
#   select-trace rebuilds with -DFD_SELECT_TRACE, prints the phase timeline of one frame's selection (tools/exp_select_trace.py), rebuilds clean
#   tune F M V.. tools/tune_kernels.sh: rebuild file F with -DM=V on the box and time the stages (e.g. tune fd_brief.cu BRIEF_PERSISTENT 0 1)
set -u
mkdir -p gpurun_out
what=${1:-tests}
case "$what" in
  tests)    python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log ;;
  fuzz)     python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/pytest_fuzz.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_fuzz.log ;;
  guard)    python tools/sanitize_paths.py > gpurun_out/sanitize_guard.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/sanitize_guard.log ;;
  bench)    tag=${2:-run}; python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_$tag.json ;;
  bench-n)  n=${2:-2}; tag=${3:-run}; python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 \
              bench.py --gpus "$n" > gpurun_out/bench_${n}gpu_$tag.json 2> gpurun_out/bench_${n}gpu_$tag.err; echo "rc=$?"; tail -c 900 gpurun_out/bench_${n}gpu_$tag.json ;;
  nccl)     python -m pytest tests/test_sharding.py tests/test_tiling.py -m gpu -x -q -k nccl > gpurun_out/pytest_nccl.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_nccl.log ;;
  launches) python bench.py --steps 2 --warmup 3 --no-extras > /dev/null 2>&1 && \
            ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
              python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_launches.log ;;
  ncu)      tag=${2:-run}; python bench.py --steps 1 --warmup 3 --no-extras > /dev/null 2>&1 && \
            ncu --set full --clock-control none --import-source on -k regex:'fast_sparse_kernel|select_kernel|brief_kernel' -c 3 -f -o gpurun_out/prof_$tag \
              python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_$tag.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_$tag.log ;;
  pipes)    make -s -C tools/microbench pipes 2>/dev/null || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/pipes tools/microbench/pipes.cu; \
            tools/microbench/pipes > gpurun_out/pipes.txt 2>&1; echo "rc=$?"; cat gpurun_out/pipes.txt ;;
  stages)   python tools/stage_times.py > gpurun_out/stage_times.json 2> gpurun_out/stage_times.err; echo "rc=$?"; cat gpurun_out/stage_times.json ;;
  tune)     shift; tools/tune_kernels.sh "$@" ;;
  one-frame) python tools/exp_single_latency.py | tee gpurun_out/single_latency.json ;;
  select-trace)
            make -s -C feature_detector_b200/csrc clean > /dev/null; make -s -j16 -C feature_detector_b200/csrc EXTRA=-DFD_SELECT_TRACE > gpurun_out/trace_build.log 2>&1 || tail -5 gpurun_out/trace_build.log
            python tools/exp_select_trace.py | tee gpurun_out/select_trace.json
            make -s -C feature_detector_b200/csrc clean > /dev/null; make -s -j16 -C feature_detector_b200/csrc > /dev/null 2>&1 ;;
  profile-r2)
            python bench.py --steps 2 --warmup 3 --no-extras > /dev/null 2>&1 && \
            ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_r2_launches.log 2>&1; echo "launches rc=$?"
            python bench.py --steps 1 --warmup 3 --no-extras > /dev/null 2>&1 && \
            ncu --set full --clock-control none --import-source on -k regex:'fast_sparse_kernel|select_kernel|brief_kernel' -s 9 -c 3 -f -o gpurun_out/prof_r2_headline python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_r2_headline.log 2>&1; echo "headline rc=$?"
            python tools/prof_r2.py all > /dev/null 2>&1 || { echo "tools/prof_r2.py failed without ncu"; exit 1; }
            ncu --set full --clock-control none --import-source on -k regex:corner_tma_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_c2 python tools/prof_r2.py c2 > gpurun_out/ncu_r2_c2.log 2>&1
            ncu --set full --clock-control none --import-source on -k regex:corner_tma_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_c3 python tools/prof_r2.py c3 > gpurun_out/ncu_r2_c3.log 2>&1
            ncu --set full --clock-control none -k regex:gather_tiles_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_gather python tools/prof_r2.py c3 > gpurun_out/ncu_r2_gather.log 2>&1
            ncu --set full --clock-control none --import-source on -k regex:'lsd_kernel|lsd_order_kernel|lsd_scatter_kernel' -s 3 -c 3 -f -o gpurun_out/prof_r2_c4 python tools/prof_r2.py c4 > gpurun_out/ncu_r2_c4.log 2>&1
            ncu --set full --clock-control none -k regex:match_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_match python tools/prof_r2.py headline > gpurun_out/ncu_r2_match.log 2>&1
            du -sh gpurun_out ;;
  *)        echo "unknown: $what"; sed -n 2,22p "$0"; exit 2 ;;
esac
