#!/bin/bash
# Round-2 run D: full-attention variants (polynomial exp2 share) + parity of the default build
mkdir -p gpurun_out
python tools/attn_bench.py > gpurun_out/attn_bench_default.json 2> gpurun_out/attn_bench.err; cat gpurun_out/attn_bench_default.json; tail -2 gpurun_out/attn_bench.err
for v in 0 2 3; do python tools/attn_bench.py --lib zoomearth_b200/_variants/libzoomvit_poly$v.so > gpurun_out/attn_bench_poly$v.json 2>> gpurun_out/attn_bench.err; cat gpurun_out/attn_bench_poly$v.json; done
python tools/attn_bench.py > gpurun_out/attn_bench_default2.json 2>> gpurun_out/attn_bench.err; cat gpurun_out/attn_bench_default2.json
timeout 600 python -m pytest tests/test_gpu_attn.py tests/test_gpu_tower.py -q -s -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/test_gpu_attn_tower.log 2>&1; echo "tests exit $?: $(tail -n 1 gpurun_out/test_gpu_attn_tower.log)"; grep -E "^(FAILED|ERROR)|^E |PARITY" gpurun_out/test_gpu_attn_tower.log | head -30
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-latency --no-sharded --no-e2e > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_d.json')); print(round(d['value']), round(d['ms_per_step'],1), d['clocks'], d['kernel_ms'])"
