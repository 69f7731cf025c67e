set -x
timeout 400 python -m pytest tests/test_gpu_live.py tests/test_gpu_at_size.py -x -q --tb=short -p no:cacheprovider --timeout 150 2>&1 | tail -3
B="timeout 150 python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 30"
E=gpurun_out/r2s26.err
: > $E
$B > gpurun_out/r2s26_f32_10M.json 2>> $E
$B --rows 1000000 > gpurun_out/r2s26_f32_1M.json 2>> $E
$B --rows 1250000 > gpurun_out/r2s26_f32_shard.json 2>> $E
$B --rows 500000 > gpurun_out/r2s26_f32_500k.json 2>> $E
$B --batch 1024 --steps 10 > gpurun_out/r2s26_f32_b1024.json 2>> $E
$B --metric l2 > gpurun_out/r2s26_f32_10M_l2.json 2>> $E
$B --dtype f16 --dim 512 --rows 6250000 --batch 4096 --steps 5 > gpurun_out/r2s26_f16_shard.json 2>> $E
tail -3 $E
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s26_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; st=d.get('search_stats',{})
        print(f.split('/')[-1][7:-5].ljust(22), round(d['value']), round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'L/step', d['gpu_launches']/d['steps'], 'ovf', d.get('overflow_rescans'), 'resc/q', round(st.get('rescored_rows_per_query',0)), 'defer/q', round(st.get('deferred_rows_per_query',0)), d['full_size_properties'].get('sampled_rows_beating_kth'))
    except Exception as e: print(f, 'ERR', e)
PY
