set -x
for pf in 0 2 4; do
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 128 --no-cpu --opt tc_prefetch_tiles=$pf > gpurun_out/x_i8_b128_pf$pf.json 2>> gpurun_out/x_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --batch 128 --no-cpu --opt tc_prefetch_tiles=$pf > gpurun_out/x_f32_b128_pf$pf.json 2>> gpurun_out/x_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --batch 256 --no-cpu --opt tc_prefetch_tiles=$pf > gpurun_out/x_f32_b256_pf$pf.json 2>> gpurun_out/x_err.log
done
timeout 300 python bench.py --steps 5 --warmup 3 --batch 256 --no-cpu --opt chunk_growth_x100=300 > gpurun_out/x_f32_b256_g3.json 2>> gpurun_out/x_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 128 --no-cpu --opt chunk_growth_x100=300 > gpurun_out/x_i8_b128_g3.json 2>> gpurun_out/x_err.log
tail -5 gpurun_out/x_err.log
