#!/bin/bash
# Round-2 run A: every GPU test file (own process each), smoke, the bench line, ncu traffic capture at HEAD.
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_gemm.py tests/test_gpu_k1.py tests/test_gpu_attn.py tests/test_gpu_tower.py tests/test_gpu_configs.py tests/test_gpu_handoff.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?; echo "== $f exit $r: $(tail -n 1 gpurun_out/$n.log)"; [ $r -ne 0 ] && { rc=1; grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/$n.log | head -20; }
done
grep -h PARITY gpurun_out/test_gpu_*.log > gpurun_out/parity.txt; cat gpurun_out/parity.txt
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['kernel_ms'], d['roofline']['frac'], d['roofline_k1']['frac'], d['cpu_baseline'], d['latency'])"
if [ "${NCU:-1}" = "1" ]; then bash tools/gpu_ncu_traffic.sh; fi
exit $rc
