#!/bin/bash
# Measurements that were prepared at the end of round 1 without GPU minutes left.  One gpurun call:
#   gpurun --timeout 900 -- 'bash tools/next_round_measurements.sh'
# Everything lands in gpurun_out/ (copy what is to be judged into profiles/).  Nothing here runs under a profiler
# except the last, explicitly separate ncu pass.
set -u
mkdir -p gpurun_out
# 1. pairs in flight: does pairs/s rise when the latency-bound chains of several pairs overlap?  (default is 1)
for n in 1 2 3 4; do
  timeout 300 python bench.py --steps 6 --warmup 3 --inflight $n --no-cpu-baseline --no-breakdown \
    > gpurun_out/r02_inflight_$n.json 2> gpurun_out/r02_inflight_$n.err || echo "inflight $n failed"
done
# 2. removeSmallSegments: scan width 1 / 4 (inside the script), prefilter off / on
FLOWB200_SEG_PREFILTER=0 timeout 200 python tests/tools/segments_time.py > gpurun_out/r02_segments_prefilter0.log 2>&1 \
  && cp gpurun_out/segments_time.json gpurun_out/r02_segments_prefilter0.json
FLOWB200_SEG_PREFILTER=1 timeout 200 python tests/tools/segments_time.py > gpurun_out/r02_segments_prefilter1.log 2>&1 \
  && cp gpurun_out/segments_time.json gpurun_out/r02_segments_prefilter1.json
# 3. Canny edge map: time and parity against cv2 at 1024x436 and 3840x2160
timeout 300 python tests/tools/edges_time.py > gpurun_out/r02_edges_time.log 2>&1
# 4. launch list of the post-processing kernels (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_post_launches.csv \
  python tests/tools/edges_time.py > gpurun_out/r02_ncu_edges.log 2>&1
