mkdir -p gpurun_out
for v in base poly44 polyaa; do
  if [ $v = base ]; then unset MMI_LIB_PATH; else export MMI_LIB_PATH=segmminterest_b200/build/variants/libmmi_$v.so; fi
  echo "== $v"
  timeout 300 python tools/attn_bench.py --dropout --bias --no-two 2>&1 | grep bwd_all
  timeout 300 python tools/attn_bench.py --bias --no-two 2>&1 | grep bwd_all
  timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py -q -x -m gpu -k "attention or attn" 2>&1 | tail -1
done
