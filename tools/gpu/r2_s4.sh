set -x
timeout 900 python -m pytest tests/test_gpu_live.py -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/r2s4_live.log 2>&1
tail -30 gpurun_out/r2s4_live.log
B="timeout 300 python bench.py --no-cpu --steps 20"
$B > gpurun_out/r2s4_f32_b256.json 2> gpurun_out/r2s4.err
$B --opt live_start_rows=32768 > gpurun_out/r2s4_f32_b256_ls32k.json 2>> gpurun_out/r2s4.err
$B --opt live_start_rows=262144 > gpurun_out/r2s4_f32_b256_ls256k.json 2>> gpurun_out/r2s4.err
$B --opt live_start_rows=1048576 > gpurun_out/r2s4_f32_b256_ls1M.json 2>> gpurun_out/r2s4.err
$B --rows 1000000 > gpurun_out/r2s4_f32_b256_1M.json 2>> gpurun_out/r2s4.err
$B --rows 1000000 --opt live_start_rows=32768 > gpurun_out/r2s4_f32_b256_1M_ls32k.json 2>> gpurun_out/r2s4.err
$B --opt live=0 > gpurun_out/r2s4_f32_b256_r1path.json 2>> gpurun_out/r2s4.err
$B --dtype i8 --batch 1024 --steps 10 > gpurun_out/r2s4_i8_b1024.json 2>> gpurun_out/r2s4.err
$B --dtype i8 --batch 1024 --steps 10 --opt live_start_rows=262144 > gpurun_out/r2s4_i8_b1024_ls256k.json 2>> gpurun_out/r2s4.err
$B --dtype i8 --batch 1024 --steps 10 --opt live=0 > gpurun_out/r2s4_i8_b1024_chunked.json 2>> gpurun_out/r2s4.err
$B --batch 1 > gpurun_out/r2s4_f32_b1.json 2>> gpurun_out/r2s4.err
$B --batch 1 --opt live=0 > gpurun_out/r2s4_f32_b1_chunked.json 2>> gpurun_out/r2s4.err
$B --batch 16 > gpurun_out/r2s4_f32_b16.json 2>> gpurun_out/r2s4.err
$B --batch 128 > gpurun_out/r2s4_f32_b128.json 2>> gpurun_out/r2s4.err
$B --batch 128 --opt live=0 > gpurun_out/r2s4_f32_b128_chunked.json 2>> gpurun_out/r2s4.err
tail -5 gpurun_out/r2s4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s4_*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']; st=d.get('search_stats',{})
        print(f.split('/')[-1][5:-5], round(d['value']), round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'L/step', d['gpu_launches']/d['steps'], 'ovf', d.get('overflow_rescans'), 'refr', round(st.get('live_refreshes_per_step',0)), 'skip', st.get('live_refresh_skips_per_step'), 'resc/q', round(st.get('rescored_rows_per_query',0)), d['full_size_properties'].get('sampled_rows_beating_kth'))
    except Exception as e: print(f, 'ERR', e)
PY
