timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fetch_threshold or batch_cost or coop or pool" 2>&1 | tail -3
for w in spheres64 spheres; do
python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-denoiser --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$w auto', j['value'], j['e2e']['value'])"
HJK_OPTIONS="coop_batch_cost=180,fetch_threshold=20" python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-denoiser --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$w old defaults', j['value'], j['e2e']['value'])"
done
