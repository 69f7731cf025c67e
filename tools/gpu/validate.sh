set -x
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/pytest_r21.log 2>&1
tail -8 gpurun_out/pytest_r21.log
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/s2_f32_b256.json 2> gpurun_out/s2_f32_b256.err; tail -2 gpurun_out/s2_f32_b256.err
timeout 600 python bench.py --dtype i8 --batch 1024 --no-cpu > gpurun_out/s2_i8_b1024.json 2> gpurun_out/s2_i8_b1024.err
timeout 600 python bench.py --dtype i8 --batch 1024 --metric dot --no-cpu > gpurun_out/s2_i8_b1024_dot.json 2>> gpurun_out/s2_i8_b1024.err
timeout 600 python bench.py --batch 1 --no-cpu > gpurun_out/s2_f32_b1.json 2> gpurun_out/s2_misc.err
timeout 600 python bench.py --batch 16 --no-cpu > gpurun_out/s2_f32_b16.json 2>> gpurun_out/s2_misc.err
timeout 600 python bench.py --rows 1000000 --no-cpu > gpurun_out/s2_f32_b256_1M.json 2>> gpurun_out/s2_misc.err
timeout 600 python bench.py --dtype i8 --batch 128 --no-cpu > gpurun_out/s2_i8_b128.json 2>> gpurun_out/s2_misc.err
timeout 600 python bench.py --dtype f16 --dim 512 --rows 6250000 --batch 4096 --no-cpu --steps 10 > gpurun_out/s2_f16_b4096_shard.json 2>> gpurun_out/s2_misc.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s2_reference.json 2> gpurun_out/s2_reference.err
cat gpurun_out/s2_*.json | cut -c1-1800
tail -3 gpurun_out/s2_misc.err
