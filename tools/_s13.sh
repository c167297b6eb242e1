mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_s13_tests.txt 2>&1
tail -4 gpurun_out/r02_s13_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_s13_smoke.txt 2>&1
tail -2 gpurun_out/r02_s13_smoke.txt
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r02_s13_bench_c2.json 2> gpurun_out/r02_s13_bench_c2.err
head -c 300 gpurun_out/r02_s13_bench_c2.json; echo
python tools/profile_step.py --table --steps 2 --warmup 2 > gpurun_out/r02_s13_table.txt 2>&1
head -24 gpurun_out/r02_s13_table.txt; tail -1 gpurun_out/r02_s13_table.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_s13_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_s13_bench_under_ncu.log 2>&1
