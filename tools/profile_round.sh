#!/bin/bash
# ncu evidence for profiles/: (1) metrics over one chunk (64 envs) of the whole pipeline, (2) launch list of a short bench.py run.
# usage (on the GPU box): bash tools/profile_round.sh <tag>   -> gpurun_out/<tag>_*.csv
set -u
TAG=${1:-r01}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed
mkdir -p gpurun_out
# profile_stereo.py brackets one steady-state chunk with cudaProfilerStart/Stop
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_chunk_metrics.csv python tools/profile_stereo.py 74 > gpurun_out/${TAG}_chunk.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_chunk_metrics.csv gpurun_out/${TAG}_chunk
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --num-envs 128 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches
