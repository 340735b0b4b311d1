#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|attn_|k1_|rmsnorm|gather|compose|cast_rows" -s 334 -c 167 --csv --log-file gpurun_out/launches_one_crop.csv python tools/one_crop.py 3 > gpurun_out/ncu_one_crop.log 2>&1; echo "ncu exit $?"
