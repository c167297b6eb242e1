mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py -q -x -m gpu -k "attention or attn" > gpurun_out/r02_s7_tests.txt 2>&1
tail -5 gpurun_out/r02_s7_tests.txt
timeout 300 python tools/attn_bench.py --dropout --bias > gpurun_out/r02_s7_attn_bench.txt 2>&1
MMI_LIB_PATH=segmminterest_b200/build/variants/libmmi_trace.so timeout 120 python tools/attn_trace_all.py 40 1 > gpurun_out/r02_s7_trace_cand.txt 2>&1
MMI_LIB_PATH=segmminterest_b200/build/variants/libmmi_trace.so timeout 120 python tools/attn_trace_all.py 500 1 > gpurun_out/r02_s7_trace_hist.txt 2>&1
