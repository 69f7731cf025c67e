# Final round-2 measurements on one B200 (gpurun): the full GPU suite, smoke, every committed bench line, ncu evidence.
set -x
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 400 > gpurun_out/r02_final_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02_final_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02_final_smoke.log
timeout 600 python bench.py > gpurun_out/r02_final_default.json 2> gpurun_out/r02_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_reference.json 2>> gpurun_out/r02_final.err
timeout 300 python bench.py --dtype i8 --batch 1024 --no-configs --steps 20 > gpurun_out/r02_final_i8_b1024.json 2>> gpurun_out/r02_final.err
B="timeout 200 python bench.py --no-cpu --no-configs --steps 20 --sustain-seconds 0"
$B --batch 1 > gpurun_out/r02_final_f32_b1.json 2>> gpurun_out/r02_final.err
$B --batch 16 > gpurun_out/r02_final_f32_b16.json 2>> gpurun_out/r02_final.err
$B --batch 128 > gpurun_out/r02_final_f32_b128.json 2>> gpurun_out/r02_final.err
$B --batch 1024 > gpurun_out/r02_final_f32_b1024.json 2>> gpurun_out/r02_final.err
$B --metric l2 > gpurun_out/r02_final_f32_b256_l2.json 2>> gpurun_out/r02_final.err
$B --dim 512 > gpurun_out/r02_final_f32_b256_d512.json 2>> gpurun_out/r02_final.err
$B --rows 1000000 > gpurun_out/r02_final_f32_b256_1M.json 2>> gpurun_out/r02_final.err
$B --rows 1250000 > gpurun_out/r02_final_f32_b256_shard.json 2>> gpurun_out/r02_final.err
$B --bitmap-density 0.1 > gpurun_out/r02_final_f32_b256_bitmap10.json 2>> gpurun_out/r02_final.err
$B --dtype i8 --batch 128 > gpurun_out/r02_final_i8_b128.json 2>> gpurun_out/r02_final.err
$B --dtype i8 --batch 256 > gpurun_out/r02_final_i8_b256.json 2>> gpurun_out/r02_final.err
$B --dtype i8 --batch 1024 --metric dot > gpurun_out/r02_final_i8_b1024_dot.json 2>> gpurun_out/r02_final.err
$B --dtype i8 --batch 1024 --bitmap-density 0.5 > gpurun_out/r02_final_i8_b1024_bitmap50.json 2>> gpurun_out/r02_final.err
$B --dtype f16 --dim 512 --rows 6250000 --batch 4096 --steps 5 > gpurun_out/r02_final_f16_b4096_shard.json 2>> gpurun_out/r02_final.err
tail -3 gpurun_out/r02_final.err
PKV_LIB_PATH=build_dbg/libpkv_timing.so timeout 120 python tools/gpu/tile_timing.py 10000000 256 2>&1 | tail -8 > gpurun_out/r02_tile_timing_f32_b256_10M.txt
# ncu: warm per-launch durations of whole searches (compare shares), then the two dominant launches in full
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum --cache-control none -k regex:"scan_|rescore|select|prep_|reset_|finalize" -s 40 -c 32 --csv --log-file gpurun_out/r02_launches_f32_b256_10M.csv python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 4 > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --cache-control none -k regex:"scan_|rescore|select|prep_|reset_|finalize" -s 40 -c 32 --csv --log-file gpurun_out/r02_launches_f32_b256_1M.csv python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 4 --rows 1000000 > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --cache-control none -k regex:"scan_|rescore|select|prep_|reset_|finalize" -s 24 -c 24 --csv --log-file gpurun_out/r02_launches_i8_b1024_10M.csv python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 4 --dtype i8 --batch 1024 > /dev/null 2>&1
timeout 400 $NCU --set full --import-source on -k regex:scan_img8 -s 3 -c 1 -o gpurun_out/r02_ncu_f32_b256_live python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 1 > /dev/null 2>&1
timeout 400 $NCU --set full --import-source on -k regex:scan_i8_ts -s 3 -c 1 -o gpurun_out/r02_ncu_i8_b1024_live python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 1 --dtype i8 --batch 1024 > /dev/null 2>&1
ls -la gpurun_out | grep r02_ | head -50
