set -x
B="timeout 300 python bench.py --no-cpu --steps 20"
for g in 150 200 300; do
$B --opt chunk_growth_x100=$g > gpurun_out/z_b256_g$g.json 2>> gpurun_out/z.err
$B --batch 1024 --opt chunk_growth_x100=$g > gpurun_out/z_b1024_g$g.json 2>> gpurun_out/z.err
done
$B --batch 1024 > gpurun_out/z_b1024_g400.json 2>> gpurun_out/z.err
$B --batch 16 --opt chunk_growth_x100=300 > gpurun_out/z_b16_g300.json 2>> gpurun_out/z.err
$B --batch 16 > gpurun_out/z_b16_gdef.json 2>> gpurun_out/z.err
$B --batch 128 --opt chunk_growth_x100=200 > gpurun_out/z_b128_g200.json 2>> gpurun_out/z.err
$B --batch 128 > gpurun_out/z_b128_g400.json 2>> gpurun_out/z.err
tail -n 5 gpurun_out/z.err
python tools/summarize.py gpurun_out/z_*.json | grep -o "^[^ ]*\|qps *[0-9]*\|scan_ms *[0-9.]*\|launches [0-9.]*" | paste - - - -
