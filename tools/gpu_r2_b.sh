#!/bin/bash
# Round-2 run B: the new resize tests, operand-dtype A/B of the bench line on ONE box, cuBLAS sustained probe, ncu traffic.
mkdir -p gpurun_out
for f in tests/test_gpu_resize.py tests/test_gpu_tower.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?; echo "== $f exit $r: $(tail -n 1 gpurun_out/$n.log)"; [ $r -ne 0 ] && grep -E "^(FAILED|ERROR)|Error|assert|^E " gpurun_out/$n.log | head -40
done
grep -h PARITY gpurun_out/test_gpu_*.log
for dt in fp16 bf16 fp16; do
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-latency --operand-dtype $dt > gpurun_out/bench_$dt.json 2> gpurun_out/bench_$dt.err; echo "bench $dt exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$dt.json')); print('$dt', round(d['value']), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d['clocks'], d['kernel_ms'], 'gemm frac', round(d['roofline']['frac'],3))"
done
python tools/peak_probe.py > gpurun_out/peak_probe.json 2>&1; cat gpurun_out/peak_probe.json
bash tools/gpu_ncu_traffic.sh
