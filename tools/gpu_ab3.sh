mkdir -p gpurun_out
python -m pytest tests/test_cuda_parity.py tests/test_cuda_scale.py -x -q -m gpu 2>&1 | tail -2
for cfg in "15 15 2 262144" "15 15 2 65536"; do
  echo "== $cfg"; python tools/phase_bench.py $cfg | cut -c1-260
  GRL_LIB_PATH=build/libgrlcuda_$1.so python tools/phase_bench.py $cfg | cut -c1-260
done
