mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py -q -x -m gpu -k "layernorm or ln" 2>&1 | tail -2
python tools/profile_step.py --table --steps 2 --warmup 2 > gpurun_out/r02_s18_table.txt 2>&1
grep -E "^ln_|total" gpurun_out/r02_s18_table.txt
