#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attn.py -q -s -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/test_gpu_attn.log 2>&1; echo "attn tests exit $?: $(grep -E 'passed|failed' gpurun_out/test_gpu_attn.log | tail -n 1)"
python tools/attn_bench.py > gpurun_out/attn_bench_i.json 2> gpurun_out/attn_bench_i.err; cut -c1-330 gpurun_out/attn_bench_i.json
python tools/attn_bench.py --lib zoomearth_b200/_variants/libzoomvit_skip16.so > gpurun_out/attn_bench_i_skip16.json 2>> gpurun_out/attn_bench_i.err; cut -c1-330 gpurun_out/attn_bench_i_skip16.json
