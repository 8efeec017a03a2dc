#!/bin/bash
# gpurun -- 'bash profiles/quick.sh <tag>': GPU tests, then a short bench (no CPU / e2e legs)
TAG=${1:-q}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; tail -8 gpurun_out/pytest_$TAG.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
    print('ms_per_step',d['ms_per_step'],'value',d['value'])
    print(json.dumps(d['roofline']['kernel_ms_per_step']))
except Exception as e:
    print('bench failed',e); print(open('gpurun_out/bench_$TAG.err').read()[-2000:])
PY
