#!/bin/bash
# ncu --set full captures of every kernel of one fit (staged schedule, one 256-frame lane) and of the persistent flow
# kernel in the default bench configuration (512 frames in 3 lanes: a 171-frame launch), plus the cloud / RTree kernels.
# Run under gpurun; reports land in gpurun_out/.  Numbers printed by bench.py under ncu are not bench values.
B="python bench.py --no-extras --no-cpu-baseline --steps 1 --warmup 1"
N="ncu --set full --clock-control none --import-source on"
AVB_FLOW=0 $N -k regex:'pose_visibility|nn_kernel|lm_prep' -s 5 -c 3 -f -o gpurun_out/r1_front $B --lanes 1 --frames 256 > gpurun_out/ncu_front.log 2>&1
AVB_FLOW=0 $N -k regex:'lm_rows|lm_gram_kernel|lm_solve' -s 45 -c 3 -f -o gpurun_out/r1_eval $B --lanes 1 --frames 256 > gpurun_out/ncu_eval.log 2>&1
AVB_FLOW=1 $N -k regex:'lm_flow' -s 3 -c 1 -f -o gpurun_out/r1_flow $B > gpurun_out/ncu_flow.log 2>&1
$N -k regex:'cloud_|rtree_' -s 4 -c 4 -f -o gpurun_out/r1_images python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_images.log 2>&1
ls -la gpurun_out/*.ncu-rep
