timeout 900 python -m pytest tests -m gpu -q > gpurun_out/test33.log 2>&1; echo "pytest exit $?" >> gpurun_out/test33.log
echo done
