set -x
timeout 300 python -m pytest tests/test_gpu_corpus.py -q --tb=short -p no:cacheprovider --timeout 100 > gpurun_out/r2s15_tests.log 2>&1
tail -15 gpurun_out/r2s15_tests.log
( time timeout 600 python bench.py > gpurun_out/r2s15_default.json 2> gpurun_out/r2s15_default.err ) 2>&1 | tail -3
tail -3 gpurun_out/r2s15_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s15_default.json').read())
r=d['roofline']
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'sustained',d['sustained'] and (round(d['sustained']['value']), d['sustained']['seconds'], d['sustained']['clocks']))
print('roofline',r['bound'],round(r['achieved']),r['peak'],round(r['frac'],3),'traffic',r['traffic'],'int8peak',r['int8_peak_measured'])
print('clocks',d['clocks'])
print('parity',d.get('parity')); print('cpu',d.get('cpu_baseline'))
print('props',d.get('full_size_properties'))
for k,v in d.get('configs',{}).items(): print(k, json.dumps(v)[:900])
PY
