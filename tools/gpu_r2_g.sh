#!/bin/bash
# Round-2 run G: window attention v2 (meta block in the TMA stage)
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_attn.py tests/test_gpu_tower.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/$n.log 2>&1
  r=$?; echo "== $f exit $r: $(grep -E 'passed|failed' gpurun_out/$n.log | tail -n 1)"; [ $r -ne 0 ] && { rc=1; grep -E "^(FAILED|ERROR)|^E |zoomvit:" gpurun_out/$n.log | head -30; }
done
if [ $rc -eq 0 ]; then
python tools/attn_bench.py > gpurun_out/attn_bench_g.json 2> gpurun_out/attn_bench_g.err; cat gpurun_out/attn_bench_g.json; tail -3 gpurun_out/attn_bench_g.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded --no-latency --no-e2e > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g.json')); print(round(d['value']), round(d['ms_per_step'],1), d['clocks'], d['kernel_ms'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_win_tc" -s 30 -c 1 -o gpurun_out/prof_attn_win_tc2 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-latency --no-sharded --images 8 > gpurun_out/ncu_win.log 2>&1; echo "ncu exit $?"
fi
exit $rc
