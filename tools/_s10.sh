mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu -k "gemm or model or dropout or train" > gpurun_out/r02_s10_tests.txt 2>&1
tail -3 gpurun_out/r02_s10_tests.txt
python tools/profile_step.py --table --steps 2 --warmup 2 > gpurun_out/r02_s10_table.txt 2>&1
grep -E "gemm_tc NT M=512000|total" gpurun_out/r02_s10_table.txt
