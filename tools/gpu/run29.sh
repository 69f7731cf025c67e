set -x
timeout 600 python -m pytest tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 600 -k "non_finite or img8" > gpurun_out/pytest_r29.log 2>&1
tail -5 gpurun_out/pytest_r29.log
B="timeout 300 python bench.py --no-cpu --steps 20"
for sg in 0 20 30 45; do
$B --opt img8_peak_sigma_x10=$sg > gpurun_out/y_b256_sg$sg.json 2>> gpurun_out/y.err
done
$B --batch 512 > gpurun_out/y_b512_img8.json 2>> gpurun_out/y.err
$B --batch 512 --opt image_mask=3 --opt use_shadow=1 > gpurun_out/y_b512_f16.json 2>> gpurun_out/y.err
$B --opt chunk_growth_x100=200 > gpurun_out/y_b256_g2.json 2>> gpurun_out/y.err
$B --opt chunk_growth_x100=800 > gpurun_out/y_b256_g8.json 2>> gpurun_out/y.err
tail -n 5 gpurun_out/y.err
python tools/summarize.py gpurun_out/y_*.json
