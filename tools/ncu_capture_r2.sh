#!/bin/bash
# Round-2 (final) captures: ncu --set full of the persistent flow kernel (fp64 default and the tensor option), of the pose /
# NN kernels, and the launch list of two default bench steps.  Run under gpurun; reports land in gpurun_out/.
# Numbers printed by bench.py under ncu are not bench values.
B="python bench.py --no-extras --no-cpu-baseline --no-parity --steps 1 --warmup 1"
N="ncu --set full --clock-control none --import-source on"
$N -k regex:'lm_flow' -s 3 -c 1 -f -o gpurun_out/r2f_flow_fp64 $B > gpurun_out/r2f_ncu_fp64.log 2>&1
$N -k regex:'lm_flow' -s 3 -c 1 -f -o gpurun_out/r2f_flow_tensor $B --jtj tensor > gpurun_out/r2f_ncu_tensor.log 2>&1
$N -k regex:'pose_visibility|nn_kernel' -s 2 -c 2 -f -o gpurun_out/r2f_nnpose $B --lanes 1 > gpurun_out/r2f_ncu_front.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2f_launch.log 2>&1
ls -la gpurun_out/*.ncu-rep
