set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_|rescore|select|prep_|reset_|finalize" -s 0 -c 40 --csv --log-file gpurun_out/r2s11_launches_1M.csv python bench.py --no-cpu --steps 1 --rows 1250000 > gpurun_out/r2s11_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_img8 -s 3 -c 1 -o gpurun_out/r2s11_live_img8 python bench.py --no-cpu --steps 1 --rows 1250000 > gpurun_out/r2s11_ncu3.log 2>&1
