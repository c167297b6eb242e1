mkdir -p gpurun_out /tmp/ncu15
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_allkeys -s 9 -c 9 -o /tmp/ncu15/allkeys -f python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/r02_s15_ncu.log 2>&1
python tools/ncu_summary.py /tmp/ncu15/allkeys.ncu-rep > gpurun_out/r02_s15_ncu_allkeys.txt 2>&1
python tools/ncu_hot_lines.py /tmp/ncu15/allkeys.ncu-rep 30 >> gpurun_out/r02_s15_ncu_allkeys.txt 2>&1
grep -A1 "^attn" gpurun_out/r02_s15_ncu_allkeys.txt | cut -c1-200
