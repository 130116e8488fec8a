run() { python bench.py --steps 4 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'dev ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), {k[12:-1]: round(v['avg_ms'],2) for k,v in d['kernels'].items()})"; }
GOOEY_B200_FRONT_PRIO=0 run noprio
run prio
GOOEY_B200_CHUNK=16384 run prio_chunk16k
GOOEY_B200_CHUNK=4096 run prio_chunk4k
