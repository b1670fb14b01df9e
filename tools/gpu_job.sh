set -x
mkdir -p gpurun_out
TAG=${1:-r1e}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
GRL_PIPE_CHUNKS=1 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_nopipe.json 2>/dev/null
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/bench_$TAG.json
python -c "
import json
for f in ['gpurun_out/bench_$TAG.json','gpurun_out/bench_${TAG}_nopipe.json']:
    d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['e2e_host_obs']['value'])
"
