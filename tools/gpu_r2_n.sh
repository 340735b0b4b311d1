#!/bin/bash
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_gemm.py tests/test_gpu_k1.py tests/test_gpu_attn.py tests/test_gpu_tower.py tests/test_gpu_configs.py tests/test_gpu_handoff.py tests/test_gpu_resize.py tests/test_gpu_ingest_plugin.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?; echo "== $f exit $r: $(grep -E 'passed|failed' gpurun_out/$n.log | tail -n 1)"; [ $r -ne 0 ] && { rc=1; grep -E "^(FAILED|ERROR)|^E |zoomvit" gpurun_out/$n.log | head -20; }
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-sharded > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n.json')); print(round(d['value']), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d['clocks'], d['kernel_ms'], d['latency'], d['roofline_attn'])"
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_attn.py tests/test_gpu_k1.py tests/test_gpu_resize.py -q -m gpu -p no:cacheprovider -k "not 5000px" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log | tail -3
exit $rc
