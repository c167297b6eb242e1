mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "gemm" > gpurun_out/r02_s12_tests.txt 2>&1
tail -4 gpurun_out/r02_s12_tests.txt
timeout 600 python tools/split_gemm_error.py > gpurun_out/r02_s12_split_gemm_error.txt 2>&1
cat gpurun_out/r02_s12_split_gemm_error.txt
MMI_FP32_TC=1 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropout.py tests/test_gpu_fullsize.py -q -m gpu -k "fp32 or float32 or strict or c3" > gpurun_out/r02_s12_model_split.txt 2>&1
tail -5 gpurun_out/r02_s12_model_split.txt
timeout 600 python - > gpurun_out/r02_s12_fp32_legs.txt 2>&1 <<'P'
import argparse, torch, json
import bench
from segmminterest_b200 import synth
wl = synth.WORKLOADS["c2"]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
table = torch.randn(wl.n_rows, wl.din, generator=g, device=dev)
a = argparse.Namespace(dropout=0.1)
for b in (64, 256):
    print(json.dumps({"batch": b, "ffma": bench.fp32_leg(a, wl, table, dev, batch=b), "split_tc": bench.fp32_leg(a, wl, table, dev, batch=b, split_tc=True)}))
P
cut -c1-400 gpurun_out/r02_s12_fp32_legs.txt
