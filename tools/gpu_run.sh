for v in rs3 rs4 rs8; do
  STAG_B200_LIB=/root/repo/variants/lib_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2>&1
done
echo done
