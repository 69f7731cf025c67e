set -x
timeout 1500 python -m pytest tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/pytest_r27.log 2>&1
tail -15 gpurun_out/pytest_r27.log
B="timeout 300 python bench.py --no-cpu --steps 20"
$B > gpurun_out/x_f32_b256_img8.json 2> gpurun_out/x.err
$B --batch 1 > gpurun_out/x_f32_b1_img8.json 2>> gpurun_out/x.err
$B --batch 16 > gpurun_out/x_f32_b16_img8.json 2>> gpurun_out/x.err
$B --batch 128 > gpurun_out/x_f32_b128_img8.json 2>> gpurun_out/x.err
$B --batch 1024 > gpurun_out/x_f32_b1024_img8.json 2>> gpurun_out/x.err
$B --metric l2 > gpurun_out/x_f32_b256_l2_img8.json 2>> gpurun_out/x.err
$B --metric l2 --opt image_mask=3 --opt use_shadow=1 > gpurun_out/x_f32_b256_l2_f16.json 2>> gpurun_out/x.err
$B --metric l2 --batch 16 > gpurun_out/x_f32_b16_l2_img8.json 2>> gpurun_out/x.err
tail -n 5 gpurun_out/x.err
python tools/summarize.py gpurun_out/x_*.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/x_launches_f32_b256.csv python bench.py --no-cpu --steps 2 --warmup 1 > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/x_launches_f32_b256.csv
