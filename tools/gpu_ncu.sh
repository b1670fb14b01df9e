#!/bin/bash
# one ncu --set full capture of the turn kernel for a shape.  usage: tools/gpu_ncu.sh TAG W H P B [gym]
TAG=$1; W=$2; H=$3; P=$4; B=$5; MODE=${6:-fused}
mkdir -p gpurun_out
if [ "$MODE" = gym ]; then TARGET="tools/gym_ncu_target.py $W $B"; else TARGET="tools/ncu_target_fused.py $W $H $P $B"; fi
python $TARGET > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:grl_turn_kernel -s 30 -c 1 -f -o gpurun_out/prof_$TAG python $TARGET > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log; ls -la gpurun_out/prof_$TAG.ncu-rep
