set -x
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 > $O/r01_final_pytest_gpu.log 2>&1
tail -4 $O/r01_final_pytest_gpu.log
timeout 600 python bench.py > $O/r01_final_f32_b256.json 2> $O/f.err
timeout 600 python bench.py --dtype i8 --batch 1024 > $O/r01_final_i8_b1024.json 2>> $O/f.err
B="timeout 300 python bench.py --no-cpu"
$B --batch 1 > $O/r01_final_f32_b1.json 2>> $O/f.err
$B --batch 16 > $O/r01_final_f32_b16.json 2>> $O/f.err
$B --batch 128 > $O/r01_final_f32_b128.json 2>> $O/f.err
$B --batch 1024 > $O/r01_final_f32_b1024.json 2>> $O/f.err
$B --metric l2 > $O/r01_final_f32_b256_l2.json 2>> $O/f.err
$B --rows 1000000 > $O/r01_final_f32_b256_1M.json 2>> $O/f.err
$B --dim 512 > $O/r01_final_f32_b256_d512.json 2>> $O/f.err
$B --dtype i8 --batch 128 > $O/r01_final_i8_b128.json 2>> $O/f.err
$B --dtype i8 --batch 256 > $O/r01_final_i8_b256.json 2>> $O/f.err
$B --dtype i8 --batch 1024 --metric dot > $O/r01_final_i8_b1024_dot.json 2>> $O/f.err
$B --dtype i8 --batch 1024 --dim 512 > $O/r01_final_i8_b1024_d512.json 2>> $O/f.err
$B --dtype f16 --dim 512 --rows 6250000 --batch 4096 --steps 10 > $O/r01_final_f16_b4096_shard.json 2>> $O/f.err
$B --bitmap-density 0.1 > $O/r01_final_f32_b256_bitmap10.json 2>> $O/f.err
$B --bitmap-density 0.5 --dtype i8 --batch 1024 > $O/r01_final_i8_b1024_bitmap50.json 2>> $O/f.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r01_final_reference.json 2>> $O/f.err
tail -n 5 $O/f.err
python tools/summarize.py $O/r01_final_*.json | grep -o "^[^ ]*\|qps *[0-9]*\|e2e *[0-9]*\|scan_ms *[0-9.]*\|frac [0-9.]*" | paste - - - - -
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 500 --csv --log-file $O/r01_launches_f32_b256.csv python bench.py --no-cpu --steps 2 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 300 --csv --log-file $O/r01_launches_i8_b1024.csv python bench.py --no-cpu --steps 2 --warmup 1 --dtype i8 --batch 1024 > /dev/null 2>&1
python tools/launch_shares.py $O/r01_launches_f32_b256.csv $O/r01_launches_i8_b1024.csv | grep -v "at::"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_img8 --launch-skip 17 --launch-count 1 -o $O/r01_prof_f32_b256_img8_main -f python bench.py --no-cpu --steps 1 --warmup 1 > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_img8 --launch-skip 2 --launch-count 1 -o $O/r01_prof_f32_b1_img8_main -f python bench.py --no-cpu --steps 1 --warmup 1 --batch 1 > $O/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_i8_ts --launch-skip 4 --launch-count 1 -o $O/r01_prof_i8_b1024_ts_main -f python bench.py --no-cpu --steps 1 --warmup 1 --dtype i8 --batch 1024 > $O/ncu3.log 2>&1
ls -la $O/*.ncu-rep
