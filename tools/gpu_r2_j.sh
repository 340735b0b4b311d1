#!/bin/bash
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_k1.py tests/test_gpu_resize.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?; echo "== $f exit $r: $(grep -E 'passed|failed' gpurun_out/$n.log | tail -n 1)"; [ $r -ne 0 ] && { rc=1; grep -E "^(FAILED|ERROR)|^E |zoomvit" gpurun_out/$n.log | head -30; }
done
grep -h PARITY gpurun_out/test_gpu_resize.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded --no-latency --no-e2e > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; python -c "
import json; d=json.load(open('gpurun_out/bench_j.json')); print(round(d['value']), round(d['ms_per_step'],1), d['clocks'], d['kernel_ms'], d['roofline_k1'])"

python tools/latency.py > gpurun_out/latency_j.json 2> gpurun_out/latency_j.err; cat gpurun_out/latency_j.json | tr -d '\n '; echo
