#!/bin/bash
# Round profile set (run under gpurun on ONE B200): bench line, ncu launch list of one training step, ncu --set full
# captures of the attention kernels and of the dominant GEMM inside the step.  Outputs land in gpurun_out/.
# Usage: bash tools/run_profiles.sh <tag>
set -u
tag=${1:-sX}
out=gpurun_out
mkdir -p $out
python bench.py --steps 8 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python tools/profile_step.py --table --steps 2 --warmup 2 > $out/${tag}_table.txt 2>&1
python tools/profile_step.py --table --steps 2 --warmup 2 --dropout 0 > $out/${tag}_table_dropout_off.txt 2>&1
# launch list of the bench command itself (serialised, cold-cache: compare shares with the bench line's kernel_breakdown)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches.csv \
    python tools/profile_step.py --steps 1 --warmup 1 > $out/${tag}_launches.log 2>&1
# attention launches of step 2 (36 per step): #37 = layer-0 history-side forward; #51..53 = layer-3 history-side dq, dkv(cand keys), dkv(history keys)
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 37 -c 1 -o $out/${tag}_attn_fwd -f \
    python tools/profile_step.py --steps 1 --warmup 1 > $out/${tag}_ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 51 -c 3 -o $out/${tag}_attn_bwd -f \
    python tools/profile_step.py --steps 1 --warmup 1 > $out/${tag}_ncu_bwd.log 2>&1
# GEMMs of step 2: the first big fused-projection GEMM (N=3072) and its neighbours
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 120 -c 4 -o $out/${tag}_gemm -f \
    python tools/profile_step.py --steps 1 --warmup 1 > $out/${tag}_ncu_gemm.log 2>&1
ls -la $out | tail -20
