set -x
timeout 900 python -m pytest tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider --timeout 120 > gpurun_out/pytest_tc.log 2>&1
tail -30 gpurun_out/pytest_tc.log
timeout 600 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 128 --no-cpu > gpurun_out/bench_i8_b128_tc.json 2> gpurun_out/bench_err_tc1.log
timeout 600 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 1024 > gpurun_out/bench_i8_b1024_tc.json 2> gpurun_out/bench_err_tc2.log
tail -3 gpurun_out/bench_err_tc1.log gpurun_out/bench_err_tc2.log
cat gpurun_out/bench_i8_b128_tc.json gpurun_out/bench_i8_b1024_tc.json
