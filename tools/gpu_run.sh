timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/test27.log 2>&1; echo "pytest exit $?" >> gpurun_out/test27.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench27.log 2>&1
STAG_NO_WIDE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench27_nowide.log 2>&1
echo done
