set -x
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_img8 --launch-skip 0 --launch-count 2 -o $O/prof_img8_tiny -f python bench.py --no-cpu --steps 1 --warmup 1 --rows 1000000 > $O/ncu_tiny.log 2>&1
tail -n 2 $O/ncu_tiny.log
