#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attn.py -q -s -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/test_gpu_attn.log 2>&1; echo "attn tests exit $?: $(grep -E 'passed|failed' gpurun_out/test_gpu_attn.log | tail -n 1)"
python tools/attn_bench.py > gpurun_out/attn_bench_h.json 2> gpurun_out/attn_bench_h.err; cat gpurun_out/attn_bench_h.json; tail -3 gpurun_out/attn_bench_h.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded --no-latency --no-e2e > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; python -c "
import json; d=json.load(open('gpurun_out/bench_h.json')); print(round(d['value']), round(d['ms_per_step'],1), d['clocks'], d['kernel_ms'])"
