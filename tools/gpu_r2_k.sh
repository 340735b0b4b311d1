#!/bin/bash
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_gemm.py tests/test_gpu_tower.py tests/test_gpu_configs.py tests/test_gpu_handoff.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?; echo "== $f exit $r: $(grep -E 'passed|failed' gpurun_out/$n.log | tail -n 1)"; [ $r -ne 0 ] && { rc=1; grep -E "^(FAILED|ERROR)|^E |zoomvit" gpurun_out/$n.log | head -30; }
done
python tools/latency.py > gpurun_out/latency_k.json 2> gpurun_out/latency_k.err; cat gpurun_out/latency_k.json | tr -d '\n '; echo
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded --no-e2e > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; python -c "
import json; d=json.load(open('gpurun_out/bench_k.json')); print(round(d['value']), round(d['ms_per_step'],1), d['clocks'], d['kernel_ms'], d['latency'])"
exit $rc
