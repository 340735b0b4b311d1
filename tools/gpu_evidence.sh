#!/bin/bash
# Round-2 evidence run on one B200 at HEAD: every GPU test file (own process each), smoke, the full bench line, the ragged
# configs, ncu launch list + traffic + full captures of the dominant kernels (FULL=k1: of the K1 kernels only),
# compute-sanitizer (SAN_TESTS: test files).  Outputs under gpurun_out/.
mkdir -p gpurun_out
rc=0
if [ "${TESTS:-files}" = "single" ]; then
  # exactly what the driver runs at round end: every GPU test in ONE process
  timeout 1500 python -m pytest tests/ -x -q -s -m gpu -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1
  r=$?; echo "== pytest tests/ -m gpu (one process) exit $r: $(grep -E 'passed|failed' gpurun_out/test_gpu_all.log | tail -n 1)"; [ $r -ne 0 ] && { rc=1; grep -E "^(FAILED|ERROR)|^E " gpurun_out/test_gpu_all.log | head -20; }
else
for f in tests/test_gpu_gemm.py tests/test_gpu_k1.py tests/test_gpu_attn.py tests/test_gpu_tower.py tests/test_gpu_configs.py tests/test_gpu_handoff.py tests/test_gpu_resize.py tests/test_gpu_ingest_plugin.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?; echo "== $f exit $r: $(grep -E 'passed|failed' gpurun_out/$n.log | tail -n 1)"; [ $r -ne 0 ] && { rc=1; grep -E "^(FAILED|ERROR)|^E " gpurun_out/$n.log | head -20; }
done
fi
grep -h PARITY gpurun_out/test_gpu_*.log | sed 's/^[.sx]*//' > gpurun_out/parity.txt; wc -l gpurun_out/parity.txt
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(round(d['value']), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d['clocks'], d['kernel_ms'], 'gemm', round(d['roofline']['frac'],3), 'k1', round(d['roofline_k1']['frac'],3), d['cpu_baseline'], d['latency'], d['sharded'])"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"; cut -c1-400 gpurun_out/bench_reference.json
python tools/bench_configs.py --config 3 4 5 > gpurun_out/configs_n1.jsonl 2> gpurun_out/configs_n1.err; echo "configs exit $?"; python -c "
import json
for l in open('gpurun_out/configs_n1.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in d if k in ('config','tokens','value','tokens_per_s','ms','tower_tflops')})"
python tools/latency.py > gpurun_out/latency.json 2> gpurun_out/latency.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|attn_|k1_|rmsnorm|gather|compose|cast_rows" -s 504 -c 168 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-latency --no-sharded --images 8 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
  bash tools/gpu_ncu_traffic.sh
  if [ "${FULL:-all}" = "all" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc" -s 133 -c 4 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-latency --no-sharded --images 8 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -s 4 -c 1 -o gpurun_out/prof_attn_tc -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-latency --no-sharded --images 8 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn_tc exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_win_tc" -s 30 -c 1 -o gpurun_out/prof_attn_win_tc -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-latency --no-sharded --images 8 > gpurun_out/ncu_win.log 2>&1; echo "ncu attn_win exit $?"
  fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_" -s 6 -c 3 -o gpurun_out/prof_k1 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-latency --no-sharded --images 8 > gpurun_out/ncu_k1.log 2>&1; echo "ncu k1 exit $?"
fi
if [ "${SAN:-1}" = "1" ]; then
  timeout 1500 compute-sanitizer --tool memcheck python -m pytest ${SAN_TESTS:-tests/test_gpu_attn.py tests/test_gpu_gemm.py tests/test_gpu_handoff.py tests/test_gpu_k1.py tests/test_gpu_resize.py} -q -m gpu -p no:cacheprovider -k "not 5000px" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log | tail -3
fi
exit $rc
