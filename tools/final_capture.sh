#!/bin/bash
# Round-end evidence on one B200: GPU tests, the default bench line, the ncu launch list of two forwards and a full
# capture of the attention kernel.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
T=${1:-r2_zz}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; tail -3 gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py > gpurun_out/${T}_bench_bf16x3.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench_bf16x3.json; tail -3 gpurun_out/${T}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_bf16x3_b32_n2048.csv python bench.py --profile > gpurun_out/${T}_ncu_launches.out 2>&1; tail -2 gpurun_out/${T}_ncu_launches.out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_attention -s 2 -c 1 -o gpurun_out/${T}_attention python tools/attn_bench.py --precision bf16x3 --iters 3 > gpurun_out/${T}_ncu_attn.out 2>&1; tail -2 gpurun_out/${T}_ncu_attn.out
ls -la gpurun_out | tail -12
