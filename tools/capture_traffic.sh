#!/bin/bash
# DRAM bytes + device time of every kernel of the steps of each BASELINE config (run on the GPU box):
#   tools/capture_traffic.sh <tag>   ->  gpurun_out/<tag>_traffic_c{1..5}.csv   (a light ncu pass: three metrics, no
# report file -- a full-metric capture of all kernels of all five configs took > 25 min and > 64 MB)
# bench.py --steps 1 --warmup 0 --no-graph --no-e2e runs three identical eager steps (launch count, timed, profiled);
# tools/ncu_traffic.py counts the steps from the once-per-step bn_finalize kernel anyway.  Only this library's kernels
# are profiled (mangled names contain the cl3d namespace): torch's input generation is skipped, not replayed.
TAG=$1
mkdir -p gpurun_out
for c in 1 2 3 4 5; do
  timeout 300 ncu --kernel-name-base mangled -k regex:cl3d \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --csv --log-file gpurun_out/${TAG}_traffic_c$c.csv \
    python bench.py --config $c --steps 1 --warmup 0 --no-graph --no-extra-configs --no-cpu-baseline --no-ref-gpu --no-e2e \
    > gpurun_out/${TAG}_traffic_c$c.log 2>&1
  grep -c "cl3d::" gpurun_out/${TAG}_traffic_c$c.csv
done
