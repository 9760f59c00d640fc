timeout 600 python -m pytest tests -m gpu -q > gpurun_out/test19.log 2>&1; echo "pytest exit $?" >> gpurun_out/test19.log
python bench.py --steps 5 --warmup 3 --mode vi --no-cpu-baseline > gpurun_out/bench19_vi.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench19.log 2>&1
echo done
