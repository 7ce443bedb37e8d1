set -x
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -120 > gpurun_out/r2/pytest1.log
tail -15 gpurun_out/r2/pytest1.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2/bench0.json 2> gpurun_out/r2/bench0.err
cat gpurun_out/r2/bench0.json
for tool in memcheck initcheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_step.py phiseg_7_5 64 2 > gpurun_out/r2/san_$tool.log 2>&1
  tail -5 gpurun_out/r2/san_$tool.log
done
# ncu --set full of the kernels furthest below their roofline (one capture per kind, last launch of each shape)
SH="64,128,128,32,32 64,128,128,192,32 64,16,16,192,192 64,4,4,192,192 64,64,64,64,64"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_halo|conv_tc' -o gpurun_out/r2/weak_stats python tools/ncu_shapes.py stats $SH > gpurun_out/r2/ncu_stats.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wgrad' -o gpurun_out/r2/weak_wgrad python tools/ncu_shapes.py wgrad $SH > gpurun_out/r2/ncu_wgrad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'norm_' -o gpurun_out/r2/weak_norm python tools/ncu_shapes.py norm 64,128,128,32,32 64,128,128,32,128 64,16,16,32,192 > gpurun_out/r2/ncu_norm.log 2>&1
python tools/bench_conv.py > gpurun_out/r2/bench_conv0.txt 2>&1
cat gpurun_out/r2/bench_conv0.txt
ls -la gpurun_out/r2
