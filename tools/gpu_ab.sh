set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_ab.log 2>&1; tail -3 gpurun_out/pytest_ab.log
python tools/phase_bench.py 15 15 2 262144 2>&1 | cut -c1-330
GRL_LIB_PATH=build/libgrlcuda_straddle_inline.so python tools/phase_bench.py 15 15 2 262144 2>&1 | cut -c1-330
python tools/phase_bench.py 15 15 2 65536 2>&1 | cut -c1-330
for i in 1 2; do
python bench.py --no-cpu-baseline > gpurun_out/bench_ab_staged$i.json 2>/dev/null
GRL_LIB_PATH=build/libgrlcuda_direct_scalars.so python bench.py --no-cpu-baseline > gpurun_out/bench_ab_direct$i.json 2>/dev/null
done
python -c "
import json
for f in ['staged1','direct1','staged2','direct2']:
    d=json.load(open('gpurun_out/bench_ab_%s.json'%f)); print(f, round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), round(d['e2e_host_obs']['value']/1e6,3))
"
python tools/gym_bench.py
