#!/bin/bash
# Round-end evidence run on one B200: every GPU test file, smoke, the bench line, the ragged-config throughputs, the ncu
# launch list of one step and full captures of the attention / K1 kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_gemm.py tests/test_gpu_k1.py tests/test_gpu_attn.py tests/test_gpu_tower.py tests/test_gpu_configs.py tests/test_gpu_handoff.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q -s -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?; echo "== $f exit $r: $(tail -n 1 gpurun_out/$n.log)"; [ $r -ne 0 ] && rc=1
done
grep -h PARITY gpurun_out/test_gpu_*.log > gpurun_out/parity.txt; cat gpurun_out/parity.txt
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['kernel_ms'], d['roofline']['frac'], d['roofline_k1']['frac'], d['cpu_baseline'])"
python tools/latency.py > gpurun_out/latency.json 2> gpurun_out/latency.err; echo "latency exit $?"
python tools/bench_configs.py --config 3 4 5 > gpurun_out/configs_n1.jsonl 2> gpurun_out/configs_n1.err; echo "configs exit $?"
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|attn_|k1_|rmsnorm|gather|compose|cast_rows" -s 501 -c 167 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -s 4 -c 1 -o gpurun_out/prof_attn_tc -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn_tc exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_window_kernel" -s 30 -c 1 -o gpurun_out/prof_attn_win -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_attn2.log 2>&1; echo "ncu attn_win exit $?"
  bash tools/gpu_ncu_k1.sh
fi
exit $rc
