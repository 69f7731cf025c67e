set -x
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider --timeout 400 > gpurun_out/r2s13_pytest.log 2>&1
tail -8 gpurun_out/r2s13_pytest.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
B="timeout 120 python bench.py --no-cpu --steps 20"
$B > gpurun_out/r2s13_f32_b256.json 2> gpurun_out/r2s13.err
$B --rows 1000000 > gpurun_out/r2s13_f32_b256_1M.json 2>> gpurun_out/r2s13.err
$B --dtype i8 --batch 1024 --steps 10 > gpurun_out/r2s13_i8_b1024.json 2>> gpurun_out/r2s13.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s13_*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']; st=d.get('search_stats',{})
        print(f.split('/')[-1][7:-5], round(d['value']), round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'L/step', d['gpu_launches']/d['steps'], 'ovf', d.get('overflow_rescans'), st)
    except Exception as e: print(f, 'ERR', e)
PY
