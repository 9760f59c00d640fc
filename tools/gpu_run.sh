timeout 600 python -m pytest tests/test_gpu_rng.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/test34a.log 2>&1; echo "pytest exit $?" >> gpurun_out/test34a.log
STAG_NB=4 timeout 600 python -m pytest tests/test_gpu_rng.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/test34b.log 2>&1; echo "pytest exit $?" >> gpurun_out/test34b.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench34a.log 2>&1
STAG_NB=4 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench34b.log 2>&1
echo done
