mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r02_s9_tests.txt 2>&1
tail -3 gpurun_out/r02_s9_tests.txt
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r02_s9_bench_c2.json 2> gpurun_out/r02_s9_bench_c2.err
python tools/profile_step.py --table --steps 2 --warmup 2 > gpurun_out/r02_s9_table.txt 2>&1
head -c 600 gpurun_out/r02_s9_bench_c2.json; head -30 gpurun_out/r02_s9_table.txt
