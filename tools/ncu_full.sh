#!/bin/bash
# ncu --set full of one launch of each back-end kernel (after the plain bench exited 0); exports the raw / source pages as
# CSV on the box and keeps only one .ncu-rep (gpurun_out/ is capped at 64 MiB).  usage: bash tools/ncu_full.sh <tag> [keep]
tag=${1:-r1_h}; keep=${2:-HatW}
for k in TomW HatW SnareW KickW; do
  rep=gpurun_out/prof_${tag}_wave_$k
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:wave_kernel.*$k -s 2 -c 1 -o $rep -f python bench.py --configs c2 --steps 1 --warmup 1 > gpurun_out/ncu_full_$k.log 2>&1
  ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw_$k.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page source --csv > gpurun_out/ncu_${tag}_source_$k.csv 2>/dev/null
  [ "$k" != "$keep" ] && rm -f $rep.ncu-rep
done
du -sh gpurun_out; ls gpurun_out
