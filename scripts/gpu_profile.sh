#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list of the same command, one --set full capture.
# Usage: scripts/gpu_profile.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -c 3000 gpurun_out/bench_${tag}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tdlo -s 3 -c 1 -o gpurun_out/prof_${tag} -f \
    python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/prof_${tag}.log 2>&1
tail -3 gpurun_out/prof_${tag}.log
