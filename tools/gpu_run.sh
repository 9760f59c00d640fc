timeout 900 python -m pytest tests -m gpu -q > gpurun_out/test30.log 2>&1; echo "pytest exit $?" >> gpurun_out/test30.log
timeout 300 python tools/bench_modes.py > gpurun_out/modes30.log 2>&1
echo done
