#!/bin/bash
# per-launch device times of one default bench step (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
AVB_FLOW=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 120 --csv --log-file gpurun_out/r1_launches_staged.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --lanes 1 --frames 256 > gpurun_out/b_ncu2.log 2>&1
