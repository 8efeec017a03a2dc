#!/bin/bash
# 8-GPU measurements of round 2 (one gpurun --gpus 8 call): weak scaling of cfg3, strong scaling at fixed
# 1e8 x 1e7, and BASELINE.json configs[4] (1e9 x 1e8).  Lines land in gpurun_out/, copies are kept in profiles/.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_r2_n8_weak.json 2> gpurun_out/bench_r2_n8_weak.err; echo weak rc=$?
$TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --scaling strong > gpurun_out/bench_r2_n8_strong.json 2> gpurun_out/bench_r2_n8_strong.err; echo strong rc=$?
$TR --master-port 29523 bench.py --gpus 8 --workload cfg4 --steps 3 --warmup 3 > gpurun_out/bench_r2_n8_cfg4.json 2> gpurun_out/bench_r2_n8_cfg4.err; echo cfg4 rc=$?
for f in weak strong cfg4; do grep -v Warn gpurun_out/bench_r2_n8_$f.err | grep -i -E "error|Traceback|rror:" | head -5; python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/bench_r2_n8_$f.json").read().strip().splitlines()[-1])
    print("$f", l["ms_per_step"], l["value"], l.get("sharded_verified"), l["config"]["pairs_per_s"], l["config"]["rows_out_total"])
except Exception as e:
    print("$f", "no line", e)
PY
done
