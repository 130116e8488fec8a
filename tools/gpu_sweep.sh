export GOOEY_B200_LIB=$PWD/libgooey_b200/lib/exp/lib_allw.so
for cfg in 32,32,32,32 32,32,16,16 32,32,8,8 32,32,32,16 32,32,16,32 32,32,8,16 32,32,32,8 16,32,32,32 32,16,32,32; do
  GOOEY_B200_WAVE_G=$cfg python bench.py --steps 3 --warmup 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('G=$cfg', 'dev ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), {k[12:-1]: round(v['avg_ms'],2) for k,v in d['kernels'].items()})"
done
