mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_ab.log 2>&1; tail -2 gpurun_out/pytest_ab.log
for cfg in "15 15 2 262144" "10 10 2 262144" "10 10 2 65536" "20 20 2 65536" "20 20 4 65536"; do
  echo "== $cfg"; python tools/phase_bench.py $cfg | cut -c1-260
  GRL_LIB_PATH=build/libgrlcuda_$1.so python tools/phase_bench.py $cfg | cut -c1-260
done
python tools/gym_prof.py 15 2>&1 | grep grl_turn | cut -c1-60,140-215; python tools/gym_prof.py 20 2>&1 | grep grl_turn | cut -c1-60,140-215
