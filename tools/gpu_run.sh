timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_spmm.py -m gpu -q > gpurun_out/test32.log 2>&1; echo "pytest exit $?" >> gpurun_out/test32.log
timeout 300 python tools/bench_gemm.py > gpurun_out/gemm32.log 2>&1
echo done
