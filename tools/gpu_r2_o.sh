#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded --no-latency > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; python -c "
import json; d=json.load(open('gpurun_out/bench_o.json')); print(round(d['value']), round(d['ms_per_step'],1), 'e2e', d['e2e'])"
