#!/bin/bash
# attention iteration loop: parity tests, ragged config throughput, one ncu capture of the tcgen05 attention kernel, quick bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attn.py tests/test_gpu_tower.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
python tools/bench_configs.py --config ${CONFIGS:-3 5} > gpurun_out/configs_n1.jsonl 2> gpurun_out/configs_n1.err; echo "configs exit $?"
python - <<P
import json
for l in open("gpurun_out/configs_n1.jsonl"):
    d = json.loads(l); print(d["workload"][:12], round(d["tokens_per_s"]), d["ms"], "attn_full", d["kernel_ms_rank0"]["attn_full"])
P
tail -3 gpurun_out/configs_n1.err
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -s 4 -c 1 -o gpurun_out/prof_attn_tc -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn_tc exit $?"
fi
python bench.py --no-cpu --no-e2e > gpurun_out/bench_quick.json
python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['kernel_ms'])"
