set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/r2/pytest3_full.log
grep -E "passed|failed|FAILED|^reproducib|^full-size|^fast |fast sampling|dp-equiv|worst relative|max \|logit|Error" gpurun_out/r2/pytest3_full.log | head -80 > gpurun_out/r2/pytest3.log
tail -60 gpurun_out/r2/pytest3_full.log >> gpurun_out/r2/pytest3.log
rm -f gpurun_out/r2/pytest3_full.log
cat gpurun_out/r2/pytest3.log | head -70
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2/bench1.json 2> gpurun_out/r2/bench1.err
cat gpurun_out/r2/bench1.json; tail -3 gpurun_out/r2/bench1.err
python tools/bench_conv.py > gpurun_out/r2/bench_conv1.txt 2>&1; cat gpurun_out/r2/bench_conv1.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2/weak python tools/ncu_shapes.py stats:64,128,128,32,32 dgrad:64,128,128,32,32 wgrad:64,128,128,32,32 stats:64,64,64,64,64 stats:64,16,16,192,192 stats:64,128,128,192,32 wgrad:64,16,16,192,192 stats:64,4,4,192,192 norm:64,128,128,32,32 > gpurun_out/r2/ncu_weak.log 2>&1
tail -3 gpurun_out/r2/ncu_weak.log
python tools/ncu_digest.py gpurun_out/r2/weak.ncu-rep stats128x32,dgrad128x32,wgrad128x32,stats64x64,stats16x192,stats128x192to32,wgrad16x192,stats4x192,act128x32,reduce128x32,apply128x32 --source 10 > gpurun_out/r2/weak_digest.txt 2>&1
ls -la gpurun_out/r2/weak.ncu-rep; rm -f gpurun_out/r2/weak.ncu-rep
head -40 gpurun_out/r2/weak_digest.txt
du -sh gpurun_out
