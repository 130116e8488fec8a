#!/bin/bash
# ncu --set full of an early launch of two back-end kernels, both .ncu-rep kept (source view).  usage: tools/ncu_two.sh tag K1 K2
tag=$1
for k in $2 $3; do
  rep=gpurun_out/prof_${tag}_wave_$k
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:wave_kernel.*$k -s 1 -c 1 -o $rep -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
