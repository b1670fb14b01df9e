#!/bin/bash
# The round's closing evidence run on one B200 (shorter than gpu_final.sh): GPU suite, both bench arms, ncu launch
# list of the bench, ncu --set full of the headline kernel (source of profiles/traffic.json).  usage: tools/gpu_final2.sh TAG
TAG=${1:-r2r}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || { tail -5 gpurun_out/${TAG}_bench.err; exit 1; }
timeout 200 python tools/phase_bench.py 20 20 2 65536 --ncu > gpurun_out/${TAG}_plain_main20.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grl_turn_kernel -s 290 -c 1 -f -o gpurun_out/prof_${TAG}_main20 \
    python tools/phase_bench.py 20 20 2 65536 --ncu > gpurun_out/${TAG}_ncu_main20.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
ls -la gpurun_out | grep ${TAG}
