python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/bench12.json 2> gpurun_out/bench12.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench12.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'],'launches',d['gpu_launches'],d['clocks'])
PY
