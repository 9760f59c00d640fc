python -m pytest tests -m gpu -q -x > gpurun_out/test11.log 2>&1; echo "pytest exit $?" >> gpurun_out/test11.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench11.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:agg_stream -s 6 -c 2 -o gpurun_out/prof11 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu11.log 2>&1
echo done
