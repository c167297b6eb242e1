#!/bin/bash
# Round-2 profile set (run under gpurun on ONE B200).  Outputs land in gpurun_out/ (kept under 64 MiB: the reports are
# summarised on the box -- tools/ncu_summary.py, plus the hottest source lines -- and only the small ones travel back).
#   launch list of the bench command, and `ncu --set full` captures of the kernels the bench line's roofline and
#   kernel_breakdown name: all-keys attention backward, attention forward, the candidate-side dq / dkv pair, the GEMM
#   instantiations, gather, LayerNorm, head/loss, AdamW.
# Usage: bash tools/run_profiles_r02.sh <tag>
set -u
tag=${1:-r02_sX}
out=gpurun_out
tmp=/tmp/ncu_$tag
mkdir -p $out $tmp
P="python tools/profile_step.py --steps 1 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_bench_under_ncu.log 2>&1
cap() {  # name, kernel regex, skip, count, keep-report(0/1), extra ncu flags
    local name=$1 rx=$2 skip=$3 cnt=$4 keep=$5; shift 5
    timeout 600 ncu --set full --clock-control none --import-source on "$@" -k "regex:$rx" -s $skip -c $cnt -o $tmp/${tag}_$name -f $P > $out/${tag}_ncu_$name.log 2>&1
    python tools/ncu_summary.py $tmp/${tag}_$name.ncu-rep > $out/${tag}_ncu_$name.txt 2>&1
    python tools/ncu_hot_lines.py $tmp/${tag}_$name.ncu-rep >> $out/${tag}_ncu_$name.txt 2>&1
    if [ "$keep" = 1 ] && [ $(stat -c %s $tmp/${tag}_$name.ncu-rep) -lt 12000000 ]; then cp $tmp/${tag}_$name.ncu-rep $out/; fi
}
# step 2 = the launches after the warm-up step's; per step: 9 all-keys backward (4 history-side + 5 candidate-side), 9 forward,
# 2 gathers, 20 + 20 LayerNorm, head fwd/bwd, loss, adamw = 64 launches, one report (summarised on the box: ~7 GPU-minutes)
cap step "(attn_|gather_l1norm|layernorm_|head_fwd|head_bwd|loss_kernel|adamw_kernel)" 64 64 0
cap gemm gemm_tc 115 36 0
du -sh $out
ls -la $out | grep $tag
MMI_LIB_PATH=segmminterest_b200/build/variants/libmmi_trace.so python tools/attn_trace_all.py 40 1 > $out/${tag}_trace_allkeys_cand.txt 2>&1
MMI_LIB_PATH=segmminterest_b200/build/variants/libmmi_trace.so python tools/attn_trace_all.py 500 1 > $out/${tag}_trace_allkeys_hist.txt 2>&1
