timeout 600 python -m pytest tests/test_gpu_spmm.py -m gpu -q -k host_buffer > gpurun_out/test31.log 2>&1; echo "pytest exit $?" >> gpurun_out/test31.log
echo done
