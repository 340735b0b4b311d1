#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attn.py -q -s -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/test_gpu_attn.log 2>&1; echo "attn tests exit $?: $(grep -E 'passed|failed' gpurun_out/test_gpu_attn.log | tail -n 1)"; grep -E "^E |zoomvit" gpurun_out/test_gpu_attn.log | head -5
python tools/win_trace.py zoomearth_b200/_variants/libzoomvit_trace.so 2>&1 | tail -17 | awk '{print $1, $8, $9, $10, $15, $16, $17, $11, $12, $13, $14}'
python tools/attn_bench.py 2>/dev/null | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded --no-latency --no-e2e > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_m.json')); print(round(d['value']), round(d['ms_per_step'],1), d['clocks'], d['kernel_ms'], d['roofline_k1']['alone'])"
