set -x
timeout 1200 python -m pytest tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider --timeout 300 -x -k "ts_" > gpurun_out/pytest_r25.log 2>&1
tail -5 gpurun_out/pytest_r25.log
B="timeout 300 python bench.py --dtype i8 --batch 1024 --no-cpu --steps 20"
for c in 1 2 3; do for g in 2 4; do
$B --opt ts_chunks=$c --opt ts_groups=$g > gpurun_out/v_c${c}_g${g}.json 2>> gpurun_out/v.err
done; done
$B --opt ts_chunks=3 --opt ts_groups=1 > gpurun_out/v_c3_g1.json 2>> gpurun_out/v.err
timeout 300 python bench.py --dtype i8 --batch 256 --no-cpu --steps 20 > gpurun_out/v_b256.json 2>> gpurun_out/v.err
timeout 300 python bench.py --dtype i8 --batch 512 --no-cpu --steps 20 > gpurun_out/v_b512.json 2>> gpurun_out/v.err
timeout 300 python bench.py --dtype i8 --batch 4096 --no-cpu --steps 10 --opt ts_groups=4 > gpurun_out/v_b4096_g4.json 2>> gpurun_out/v.err
tail -n 3 gpurun_out/v.err
python tools/summarize.py gpurun_out/v_*.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_i8_ts --launch-skip 4 --launch-count 1 -o gpurun_out/r01_prof_i8_ts_c3g4 -f python bench.py --dtype i8 --batch 1024 --no-cpu --steps 1 --warmup 1 --opt ts_groups=4 > gpurun_out/ncu_ts.log 2>&1
