set -x
mkdir -p gpurun_out
for W in ${1:-20 15}; do
python tools/gym_ncu_target.py $W > gpurun_out/gym_plain_$W.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grl_turn_kernel -s 30 -c 1 -f -o gpurun_out/prof_gym$W python tools/gym_ncu_target.py $W > gpurun_out/ncu_gym_$W.log 2>&1
done
ls -la gpurun_out | tail -5
