set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --batch 8 --rows 2000000 > gpurun_out/bench_f32_b8_2M.json 2> gpurun_out/bench_err1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f32_b256.json 2> gpurun_out/bench_err2.log
timeout 600 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 64 > gpurun_out/bench_i8_b64.json 2> gpurun_out/bench_err3.log
tail -5 gpurun_out/pytest.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_*.json
