set -x
TAG=${1:-r1f}
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
python bench.py --steps 20 --warmup 3 --quick > gpurun_out/plainq_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grl_turn_kernel -s 30 -c 2 -f -o gpurun_out/prof_$TAG python bench.py --steps 20 --warmup 3 --quick > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -5
