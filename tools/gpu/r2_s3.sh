set -x
timeout 900 python -m pytest tests/test_gpu_live.py -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/r2s3_live.log 2>&1
tail -40 gpurun_out/r2s3_live.log
B="timeout 300 python bench.py --no-cpu --steps 20"
$B > gpurun_out/r2s3_f32_b256.json 2> gpurun_out/r2s3.err
$B --rows 1000000 > gpurun_out/r2s3_f32_b256_1M.json 2>> gpurun_out/r2s3.err
$B --opt live=0 --opt img8_fused=0 > gpurun_out/r2s3_f32_b256_r1path.json 2>> gpurun_out/r2s3.err
$B --opt live=0 > gpurun_out/r2s3_f32_b256_chunkedfused.json 2>> gpurun_out/r2s3.err
$B --dtype i8 --batch 1024 --steps 10 > gpurun_out/r2s3_i8_b1024.json 2>> gpurun_out/r2s3.err
$B --dtype i8 --batch 1024 --steps 10 --opt live=0 > gpurun_out/r2s3_i8_b1024_chunked.json 2>> gpurun_out/r2s3.err
$B --dtype i8 --batch 1024 --steps 10 --opt live_refresh=64 > gpurun_out/r2s3_i8_b1024_refresh64.json 2>> gpurun_out/r2s3.err
$B --batch 1 > gpurun_out/r2s3_f32_b1.json 2>> gpurun_out/r2s3.err
tail -5 gpurun_out/r2s3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s3_*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],3), 'kernel_ms', round(r['kernel_ms_per_step'],3), 'launches/step', d['gpu_launches']/d['steps'], d.get('overflow_rescans'), d.get('search_stats'), d['full_size_properties'].get('sampled_rows_beating_kth'))
    except Exception as e: print(f, 'ERR', e)
PY
