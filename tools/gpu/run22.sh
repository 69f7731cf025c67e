set -x
timeout 900 python -m pytest tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/pytest_r22.log 2>&1
tail -15 gpurun_out/pytest_r22.log
B="timeout 300 python bench.py --dtype i8 --batch 1024 --no-cpu --steps 20"
$B > gpurun_out/ts_g2.json 2> gpurun_out/ts.err
$B --opt ts_groups=1 > gpurun_out/ts_g1.json 2>> gpurun_out/ts.err
$B --opt ts_groups=4 > gpurun_out/ts_g4.json 2>> gpurun_out/ts.err
$B --opt tc_ts=0 > gpurun_out/ts_off.json 2>> gpurun_out/ts.err
$B --opt ts_stages=12 > gpurun_out/ts_g2_s12.json 2>> gpurun_out/ts.err
timeout 300 python bench.py --dtype i8 --batch 256 --no-cpu --steps 20 > gpurun_out/ts_b256.json 2>> gpurun_out/ts.err
timeout 300 python bench.py --dtype i8 --batch 256 --no-cpu --steps 20 --opt tc_ts=0 > gpurun_out/ts_b256_off.json 2>> gpurun_out/ts.err
timeout 300 python bench.py --dtype i8 --batch 512 --no-cpu --steps 20 > gpurun_out/ts_b512.json 2>> gpurun_out/ts.err
tail -3 gpurun_out/ts.err
python tools/summarize.py gpurun_out/ts_*.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_i8_ts --launch-skip 8 --launch-count 2 -o gpurun_out/r01_prof_i8_ts -f python bench.py --dtype i8 --batch 1024 --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_ts.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_i8_tc2 --launch-skip 16 --launch-count 1 -o gpurun_out/r01_prof_i8_tc2 -f python bench.py --dtype i8 --batch 1024 --no-cpu --steps 1 --warmup 1 --opt tc_ts=0 > gpurun_out/ncu_tc2.log 2>&1
tail -3 gpurun_out/ncu_ts.log gpurun_out/ncu_tc2.log
ls -la gpurun_out/*.ncu-rep
