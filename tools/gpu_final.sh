#!/bin/bash
# The round's evidence run on one B200: bench (both arms), ncu launch list of the bench, ncu --set full of the final
# kernels.  usage: tools/gpu_final.sh TAG
TAG=${1:-r2g}
mkdir -p gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || { tail -5 gpurun_out/${TAG}_bench.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
# the headline kernel exactly as bench.py runs it (replayed moves, turn 200): skip the fast-forward and recording launches
python tools/phase_bench.py 20 20 2 65536 --ncu > gpurun_out/${TAG}_plain_main20.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grl_turn_kernel -s 290 -c 1 -f -o gpurun_out/prof_${TAG}_main20 \
    python tools/phase_bench.py 20 20 2 65536 --ncu > gpurun_out/${TAG}_ncu_main20.log 2>&1
python tools/phase_bench.py 15 15 2 262144 --ncu > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grl_turn_kernel -s 290 -c 1 -f -o gpurun_out/prof_${TAG}_main15 \
    python tools/phase_bench.py 15 15 2 262144 --ncu > gpurun_out/${TAG}_ncu_main15.log 2>&1
tools/gpu_ncu.sh ${TAG}_gym15 15 15 2 65536 gym
tools/gpu_ncu.sh ${TAG}_gym20 20 20 2 65536 gym
ls -la gpurun_out | grep ${TAG}
