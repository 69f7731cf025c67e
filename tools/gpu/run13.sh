set -x
timeout 600 python -m pytest tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider --timeout 180 -x > gpurun_out/pytest_pf.log 2>&1
tail -4 gpurun_out/pytest_pf.log
for pf in 0 1 2 4; do
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 256 --no-cpu --opt tc_prefetch_tiles=$pf > gpurun_out/u_i8_b256_pf$pf.json 2>> gpurun_out/u_err.log
done
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 1024 --no-cpu --opt tc_prefetch_tiles=2 > gpurun_out/u_i8_b1024_pf2.json 2>> gpurun_out/u_err.log
for g in 250 300; do
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 1024 --no-cpu --opt chunk_growth_x100=$g > gpurun_out/u_i8_b1024_g$g.json 2>> gpurun_out/u_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --batch 256 --no-cpu --opt chunk_growth_x100=$g > gpurun_out/u_f32_b256_g$g.json 2>> gpurun_out/u_err.log
done
tail -3 gpurun_out/u_err.log
