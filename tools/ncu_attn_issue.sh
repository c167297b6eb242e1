#!/bin/bash
# issue-slot / stall-reason metrics of the attention kernels at the c2 history-side shapes (raw ncu CSV to gpurun_out/)
ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_misc_per_issue_active.ratio,smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_selected_per_issue_active.ratio \
  --clock-control none -k regex:attn_ -c 4 --csv --log-file gpurun_out/${1:-attn_issue}.csv python tools/attn_bench.py --iters 1 --sides usr > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/${1:-attn_issue}.csv") if not l.startswith("=="))]
h=rows[0]
ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
iid = h.index("ID")
cur=None
for r in rows[1:]:
    if r[iid]!=cur:
        cur=r[iid]; print(r[iid], r[ik][:48])
    print("    %-70s %s" % (r[im].replace("smsp__average_warps_issue_stalled_","stall_").replace("_per_issue_active.ratio",""), r[iv]))
PY
