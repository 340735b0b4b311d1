#!/bin/bash
# smoke + bench + ncu launch list + one full ncu capture of the top GEMM; outputs under gpurun_out/
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${STEPS:-3} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 395 -c 520 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 4 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images 8 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
fi
