#!/bin/bash
# Multi-GPU run (N=2 by default): fused-gather correctness logs, the bench line at N ranks (double-buffered gather + configs[3]
# sub-record); TESTS=1 also runs two single-GPU test files (costs N x box time: leave it off at N = 8).
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py > gpurun_out/multi_gpu_check_n$N.log 2>&1; echo "multi_gpu_check exit $?"
grep -E "^rank" gpurun_out/multi_gpu_check_n$N.log | sort; tail -5 gpurun_out/multi_gpu_check_n$N.log | grep -v "^rank"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"; tail -5 gpurun_out/bench_n$N.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],1), 'e2e', d['e2e'] and round(d['e2e']['value']), d['clocks'], d.get('gather_check'), d['sharded'], d['kernel_ms'])"
[ "${TESTS:-0}" = "1" ] && CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_ingest_plugin.py tests/test_gpu_configs.py -q -s -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/test_gpu_misc.log 2>&1; echo "tests exit $?: $(tail -n 1 gpurun_out/test_gpu_misc.log)"; grep -E "^(FAILED|ERROR)|^E " gpurun_out/test_gpu_misc.log | head -30
