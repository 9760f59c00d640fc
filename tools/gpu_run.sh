timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/test17.log 2>&1; echo "pytest exit $?" >> gpurun_out/test17.log
echo done
