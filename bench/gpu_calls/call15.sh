python -m pytest tests -m gpu -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench11.json 2> gpurun_out/bench11.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench11.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'],'launches',d['gpu_launches'],d['clocks'])
for r in d['c5_sweep_kernel_only']: print(r)
PY
