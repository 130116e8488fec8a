#!/bin/bash
# One GPU-box pass for a round: parity tests, plain bench (both arms), ncu launch list, then ncu --set full of the
# back-end kernels through tools/ncu_full.sh (CSV pages exported on the box; gpurun_out/ is capped at 64 MiB).
# usage: bash tools/profile_round.sh <tag>      (outputs under gpurun_out/)
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_$tag.log 2>&1; tail -2 gpurun_out/gpu_tests_$tag.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err && cut -c1-300 gpurun_out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list_$tag.log 2>&1
bash tools/ncu_full.sh $tag TomW
