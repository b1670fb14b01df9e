set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r1d.log
python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_r1d.json 2> gpurun_out/bench_ref_r1d.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:grl_turn_kernel -s 30 -c 2 -f -o gpurun_out/prof_r1d python bench.py --steps 20 --warmup 3 --quick > gpurun_out/ncu_full_r1d.log 2>&1
tail -3 gpurun_out/pytest_gpu_r1d.log
cat gpurun_out/bench_r1d.json
