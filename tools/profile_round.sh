#!/bin/bash
# One GPU-box pass for a round: parity tests, plain bench (both arms, every config), ncu launch list of the C2 step, then
# ncu --set full of the back-end kernels through tools/ncu_full.sh (CSV pages exported on the box; gpurun_out/ is capped at
# 64 MiB) and of the engine-level kernels (effect mixer, bass, granulator).
# usage: bash tools/profile_round.sh <tag>      (outputs under gpurun_out/)
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_$tag.log 2>&1; tail -2 gpurun_out/gpu_tests_$tag.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err && cut -c1-300 gpurun_out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --configs c2 --steps 2 --warmup 1 > gpurun_out/ncu_list_$tag.log 2>&1
bash tools/ncu_full.sh $tag none
# engine-level kernels (one long piece each: the first ten pieces of a bounce are the short lead-in)
for k in chain_fast_kernel bass_wave_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 11 -c 1 -o gpurun_out/prof_${tag}_$k -f python tools/engine_scale.py --engines 2048 --bars 2 --fx --check 0 --reps 1 > gpurun_out/ncu_full_$k.log 2>&1
  ncu -i gpurun_out/prof_${tag}_$k.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw_$k.csv 2>/dev/null
  python profiles/ncu_hot_lines.py gpurun_out/prof_${tag}_$k.ncu-rep 2>/dev/null | head -45 > gpurun_out/hot_lines_${tag}_$k.txt
  rm -f gpurun_out/prof_${tag}_$k.ncu-rep          # gpurun_out/ is capped at 64 MiB: keep the text, not the reports
done
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gran_wave_kernel -s 11 -c 1 -o gpurun_out/prof_${tag}_gran -f python tools/gran_scale.py --engines 1600 --seconds 4 --check 0 > gpurun_out/ncu_full_gran.log 2>&1
ncu -i gpurun_out/prof_${tag}_gran.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw_gran_wave_kernel.csv 2>/dev/null
python profiles/ncu_hot_lines.py gpurun_out/prof_${tag}_gran.ncu-rep 2>/dev/null | head -45 > gpurun_out/hot_lines_${tag}_gran_wave_kernel.txt
rm -f gpurun_out/prof_${tag}_gran.ncu-rep
du -sh gpurun_out
