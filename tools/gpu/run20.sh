set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider --timeout 300 -k "concurrent or more_queries or wide_rows" > gpurun_out/pytest_r20.log 2>&1
tail -5 gpurun_out/pytest_r20.log
python tools/concurrent_bench.py --dtype i8 --rows 10000000 > gpurun_out/concurrent_i8_10M.jsonl 2> gpurun_out/conc_err.log
python tools/concurrent_bench.py --dtype f32 --rows 10000000 > gpurun_out/concurrent_f32_10M.jsonl 2>> gpurun_out/conc_err.log
cat gpurun_out/concurrent_*.jsonl; tail -3 gpurun_out/conc_err.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json
