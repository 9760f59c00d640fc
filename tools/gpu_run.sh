python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain28.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu28a.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain28b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:agg_stream -s 18 -c 3 -o gpurun_out/r01_agg_stream python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu28b.log 2>&1
timeout 300 python tools/bench_modes.py > gpurun_out/modes28.log 2>&1
echo done
