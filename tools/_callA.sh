mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/a_tests.log
for S in 1 2 4; do
  EPRECON_STREAMS=$S EPRECON_BENCH_SKIP_CPU=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_s$S.json 2> gpurun_out/a_bench_s$S.err; echo "bench S=$S exit $?"
done
timeout 300 python tools/host_profile.py > gpurun_out/a_hostprof.txt 2>&1
EPRECON_STREAMS=1 EPRECON_BENCH_SKIP_CPU=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/a_ncu_bench.log 2>&1; echo "ncu exit $?"
tail -5 gpurun_out/a_tests.log; cat gpurun_out/a_bench_s*.json | cut -c1-400; tail -3 gpurun_out/a_bench_s4.err
