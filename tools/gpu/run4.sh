set -x
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 180 > gpurun_out/pytest_tc2.log 2>&1
tail -40 gpurun_out/pytest_tc2.log
timeout 600 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 128 --no-cpu > gpurun_out/bench_i8_b128_tc2.json 2> gpurun_out/bench_err_tc3.log
timeout 600 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 1024 --no-cpu > gpurun_out/bench_i8_b1024_tc2.json 2> gpurun_out/bench_err_tc4.log
timeout 600 python bench.py --steps 5 --warmup 3 --batch 256 > gpurun_out/bench_f32_b256_tc.json 2> gpurun_out/bench_err_tc5.log
timeout 600 python bench.py --steps 5 --warmup 3 --batch 128 --no-cpu > gpurun_out/bench_f32_b128_tc.json 2> gpurun_out/bench_err_tc6.log
for f in gpurun_out/bench_err_tc3.log gpurun_out/bench_err_tc4.log gpurun_out/bench_err_tc5.log gpurun_out/bench_err_tc6.log; do tail -n 3 $f; done
