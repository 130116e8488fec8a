#!/bin/bash
# One GPU-box pass for a round: parity tests, plain bench (both arms), ncu launch list, ncu --set full of the back-end kernels.
# usage: bash tools/profile_round.sh <tag>      (outputs under gpurun_out/)
tag=${1:-r1_h}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_$tag.log 2>&1; tail -2 gpurun_out/gpu_tests_$tag.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err && cat gpurun_out/bench_$tag.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list_$tag.log 2>&1
for k in TomW HatW SnareW KickW; do
  ncu --set full --clock-control none --import-source on -k regex:wave_kernel.*$k -s 6 -c 1 -o gpurun_out/prof_${tag}_wave_$k -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full_$k.log 2>&1
done
ls -la gpurun_out | tail -12
