python -m pytest tests -m gpu -q > gpurun_out/test5.log 2>&1; echo "pytest exit $?" >> gpurun_out/test5.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 9 -c 2 -o gpurun_out/prof5 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu5.log 2>&1
echo done
