#!/bin/bash
# round-2 GPU job: parity subset + per-shape timings.  usage: tools/gpu_r2.sh TAG [pytest-k-expression]
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${2:+-k "$2"} 2>&1 | tail -8
for shape in "20 20 2 65536" "15 15 2 262144" "15 15 2 65536" "10 10 2 65536" "10 10 2 262144" "20 20 4 65536"; do
  python tools/phase_bench.py $shape 2>&1 | tail -1 | tee -a gpurun_out/phase_$TAG.jsonl
done
