timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/test23.log 2>&1; echo "pytest exit $?" >> gpurun_out/test23.log
echo done
