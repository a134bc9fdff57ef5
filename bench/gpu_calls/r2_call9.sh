#!/bin/bash
# round 2, call 9: per-stream-group filters (FFMA2 row-pair kernel with per-item tables, tensor kernel with per-stream tap cores)
timeout 900 python -m pytest tests/test_gpu_groups.py -x -q > gpurun_out/r2c9_pytest.txt 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2c9_pytest.txt
( time timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-sweep --sustain-seconds 0 --cpu-seconds 2 ) > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2c9_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c9_bench.json').read().strip().splitlines()[0])
for k,v in d["configs"].items():
    print(k, v["value"], "e2e", v["e2e"]["value"], v["roofline"]["kernel"][:40], v["roofline"].get("launches_per_step"))
PY
