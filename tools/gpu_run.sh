timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q > gpurun_out/test26.log 2>&1; echo "pytest exit $?" >> gpurun_out/test26.log
echo done
