#!/bin/bash
# Round-2 run E: programmatic dependent launch - all GPU tests, latency with / without PDL, quick bench
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_gemm.py tests/test_gpu_k1.py tests/test_gpu_attn.py tests/test_gpu_tower.py tests/test_gpu_configs.py tests/test_gpu_handoff.py tests/test_gpu_resize.py tests/test_gpu_ingest_plugin.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?; echo "== $f exit $r: $(grep -E 'passed|failed' gpurun_out/$n.log | tail -n 1)"; [ $r -ne 0 ] && { rc=1; grep -E "^(FAILED|ERROR)|^E " gpurun_out/$n.log | head -20; }
done
python tools/latency.py > gpurun_out/latency_pdl.json 2> gpurun_out/latency.err; echo "latency pdl:"; cat gpurun_out/latency_pdl.json | tr -d '\n'; echo
python tools/latency.py --lib zoomearth_b200/_variants/libzoomvit_nopdl.so > gpurun_out/latency_nopdl.json 2>> gpurun_out/latency.err; echo "latency nopdl:"; cat gpurun_out/latency_nopdl.json | tr -d '\n'; echo
python tools/latency.py > gpurun_out/latency_pdl2.json 2>> gpurun_out/latency.err; echo "latency pdl (2nd):"; cat gpurun_out/latency_pdl2.json | tr -d '\n'; echo
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_e.json')); print(round(d['value']), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d['clocks'], d['kernel_ms'], d['latency'])"
exit $rc
