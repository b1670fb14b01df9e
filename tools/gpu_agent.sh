#!/bin/bash
# The in-step random agent on a B200: GPU suite, the vector-env figures with the agent as a launch of its own and drawn
# inside the step (bench.gym_contract_rate at 15x15 / 10x10 / 20x20), the gym step with given actions (regression check),
# the pool bench.  usage: tools/gpu_agent.sh TAG
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python - > gpurun_out/${TAG}_gym_env.jsonl 2> gpurun_out/${TAG}_gym_env.err <<'P'
import json, bench
for board in (15, 10, 20):
    print(json.dumps(bench.gym_contract_rate(65536, board, 60)), flush=True)
P
cat gpurun_out/${TAG}_gym_env.jsonl | cut -c1-330
for s in "20 20" "15 15" "10 10"; do timeout 200 python tools/phase_bench.py $s 2 65536 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config'], 'fused', d['ms_per_launch']['fused(obs+mask+reward+done)'], 'gym', d.get('gym_step_ms'))"; done | tee gpurun_out/${TAG}_phase.txt
timeout 300 python tools/pool_bench.py > gpurun_out/${TAG}_pool_bench.jsonl 2> gpurun_out/${TAG}_pool.err; cut -c1-200 gpurun_out/${TAG}_pool_bench.jsonl
