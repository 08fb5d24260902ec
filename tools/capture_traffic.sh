#!/bin/bash
# one `ncu --set full` report of ONE eager step per BASELINE config (run on the GPU box):  tools/capture_traffic.sh <tag>
# skips the warm-up launches of bench.py (--steps 1 --warmup 0 --no-graph: first the launch-count step, then the timed
# step, then the profile step: three identical steps; the report holds all three -> steps=3 in tools/ncu_traffic.py)
TAG=$1
mkdir -p gpurun_out
for c in 1 2 3 4 5; do
  timeout 900 ncu --set full --clock-control none -o gpurun_out/${TAG}_traffic_c$c -f \
    python bench.py --config $c --steps 1 --warmup 0 --no-graph --no-extra-configs --no-cpu-baseline --no-ref-gpu \
    > gpurun_out/${TAG}_traffic_c$c.log 2>&1
  tail -1 gpurun_out/${TAG}_traffic_c$c.log | cut -c1-200
done
