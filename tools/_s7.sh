mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py -q -x -m gpu -k "attention or attn" > gpurun_out/r02_s7_tests.txt 2>&1
tail -3 gpurun_out/r02_s7_tests.txt
timeout 300 python tools/attn_bench.py --dropout --bias --no-two > gpurun_out/r02_s7_attn_bench.txt 2>&1
cat gpurun_out/r02_s7_attn_bench.txt
