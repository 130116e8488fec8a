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
bash tools/ncu_full.sh $tag TomW
# warm-cache view of the dominant kernel's DRAM traffic (is the plane traffic an artefact of ncu's cache flush?)
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --kernel-name-base demangled -k regex:wave_kernel -s 8 -c 8 --csv --log-file gpurun_out/ncu_${tag}_warm_dram.csv python bench.py --configs c2 --steps 1 --warmup 1 > /dev/null 2>&1
# engine-level kernels (one long piece each: the first ten pieces of a bounce are the short lead-in)
for k in chain_fast_kernel bass_wave_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 11 -c 1 -o gpurun_out/prof_${tag}_$k -f python tools/engine_scale.py --engines 2048 --bars 2 --fx --check 0 --reps 1 > gpurun_out/ncu_full_$k.log 2>&1
  ncu -i gpurun_out/prof_${tag}_$k.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw_$k.csv 2>/dev/null
done
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gran_wave_kernel -s 11 -c 1 -o gpurun_out/prof_${tag}_gran -f python tools/gran_scale.py --engines 1600 --seconds 4 --check 0 > gpurun_out/ncu_full_gran.log 2>&1
ncu -i gpurun_out/prof_${tag}_gran.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw_gran_wave_kernel.csv 2>/dev/null
COUNTS=1024 GOOEY_B200_COOP_ABOVE=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:coop_kernel -s 3 -c 1 -o gpurun_out/prof_${tag}_coop -f python tools/type_scaling.py tom > gpurun_out/ncu_full_coop.log 2>&1
ncu -i gpurun_out/prof_${tag}_coop.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw_coop_kernel.csv 2>/dev/null
rm -f gpurun_out/prof_${tag}_coop.ncu-rep
du -sh gpurun_out
