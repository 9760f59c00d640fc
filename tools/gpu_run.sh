python -m pytest tests/test_gpu_rng.py tests/test_gpu_spmm.py -m gpu -q -x > gpurun_out/test16.log 2>&1; echo "pytest exit $?" >> gpurun_out/test16.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench16.log 2>&1
STAG_B200_LIB=/root/repo/variants/lib_u2s4.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench16_u2s4.log 2>&1
echo done
