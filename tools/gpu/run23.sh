set -x
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/pytest_r23.log 2>&1
tail -5 gpurun_out/pytest_r23.log
B="timeout 300 python bench.py --dtype i8 --batch 1024 --no-cpu --steps 20"
$B > gpurun_out/u_ts_g2.json 2> gpurun_out/u.err
$B --opt ts_groups=1 > gpurun_out/u_ts_g1.json 2>> gpurun_out/u.err
$B --opt ts_groups=4 > gpurun_out/u_ts_g4.json 2>> gpurun_out/u.err
$B --opt tc_ts=0 > gpurun_out/u_ts_off.json 2>> gpurun_out/u.err
timeout 300 python bench.py --dtype i8 --batch 256 --no-cpu --steps 20 > gpurun_out/u_b256.json 2>> gpurun_out/u.err
timeout 300 python bench.py --dtype i8 --batch 128 --no-cpu --steps 20 > gpurun_out/u_b128.json 2>> gpurun_out/u.err
timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/u_f32_b256.json 2>> gpurun_out/u.err
timeout 300 python bench.py --no-cpu --steps 20 --batch 128 > gpurun_out/u_f32_b128.json 2>> gpurun_out/u.err
timeout 300 python bench.py --no-cpu --steps 20 --batch 1024 > gpurun_out/u_f32_b1024.json 2>> gpurun_out/u.err
timeout 300 python bench.py --dtype f16 --dim 512 --rows 6250000 --batch 4096 --no-cpu --steps 10 > gpurun_out/u_f16_b4096.json 2>> gpurun_out/u.err
tail -3 gpurun_out/u.err
python tools/summarize.py gpurun_out/u_*.json
