python -m pytest tests -m gpu -q -x > gpurun_out/test13.log 2>&1; echo "pytest exit $?" >> gpurun_out/test13.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench13.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke13.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke13.log
echo done
