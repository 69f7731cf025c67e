set -x
timeout 1500 python -m pytest tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest_r26.log 2>&1
tail -25 gpurun_out/pytest_r26.log
B="timeout 300 python bench.py --no-cpu --steps 20"
$B > gpurun_out/w_f32_b256_img8.json 2> gpurun_out/w.err
$B --opt image_mask=3 --opt use_shadow=1 > gpurun_out/w_f32_b256_f16.json 2>> gpurun_out/w.err
$B --batch 1 > gpurun_out/w_f32_b1_img8.json 2>> gpurun_out/w.err
$B --batch 16 > gpurun_out/w_f32_b16_img8.json 2>> gpurun_out/w.err
$B --batch 128 > gpurun_out/w_f32_b128_img8.json 2>> gpurun_out/w.err
$B --batch 1024 > gpurun_out/w_f32_b1024_img8.json 2>> gpurun_out/w.err
$B --metric l2 > gpurun_out/w_f32_b256_l2_img8.json 2>> gpurun_out/w.err
timeout 300 python bench.py --rows 1000000 --steps 20 > gpurun_out/w_f32_b256_1M_img8.json 2>> gpurun_out/w.err
tail -n 5 gpurun_out/w.err
python tools/summarize.py gpurun_out/w_*.json
grep -o '"parity": {[^}]*}' gpurun_out/w_f32_b256_1M_img8.json
