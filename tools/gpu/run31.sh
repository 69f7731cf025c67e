set -x
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 > gpurun_out/pytest_r31.log 2>&1
tail -15 gpurun_out/pytest_r31.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 400 --csv --log-file gpurun_out/r01_launches_f32_b256_img8.csv python bench.py --no-cpu --steps 2 --warmup 1 > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/r01_launches_f32_b256_img8.csv | grep -v "at::"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_img8 --launch-skip 9 --launch-count 1 -o gpurun_out/r01_prof_f32_b256_img8 -f python bench.py --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_img8.log 2>&1
tail -n 2 gpurun_out/ncu_img8.log
