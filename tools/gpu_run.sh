timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -k c2 > gpurun_out/test29.log 2>&1; echo "pytest exit $?" >> gpurun_out/test29.log
python bench.py > gpurun_out/bench29.log 2>&1
python bench.py --impl reference > gpurun_out/bench29_ref.log 2>&1
echo done
