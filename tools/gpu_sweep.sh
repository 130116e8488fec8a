python -m pytest tests/test_voices_gpu.py tests/test_golden_gpu.py tests/test_edges_gpu.py -m gpu -q 2>&1 | tail -2
python bench.py --steps 4 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('main', 'dev ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), {k[12:-1]: round(v['avg_ms'],2) for k,v in d['kernels'].items()})"
python tools/type_scaling.py tom 2>&1 | grep " 1024 "
