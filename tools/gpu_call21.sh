set -x
mkdir -p gpurun_out/r2
# (a) launch list of one eager single-stream training step, NVTX-labelled
timeout 420 ncu --profile-from-start off --nvtx --print-nvtx-rename kernel --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/r2/launches_r02.csv python tools/ncu_step.py --out gpurun_out/r2/step_list_r02.json > gpurun_out/r2/ncu_step.log 2>&1
python tools/launch_table.py gpurun_out/r2/launches_r02.csv 70 > gpurun_out/r2/launches_r02.txt 2>&1; head -30 gpurun_out/r2/launches_r02.txt
gzip -f gpurun_out/r2/launches_r02.csv
# (b) CUPTI timeline of one graph replay
timeout 200 python tools/timeline.py --out gpurun_out/r2/timeline_r02.json > gpurun_out/r2/timeline.log 2>&1
gzip -f gpurun_out/r2/timeline_r02_chrome.json; rm -f gpurun_out/r2/timeline_r02.json
python tools/timeline_report.py gpurun_out/r2/timeline_r02_chrome.json.gz > gpurun_out/r2/timeline_r02.txt 2>&1; head -45 gpurun_out/r2/timeline_r02.txt
# (c) the dominant kernel, full counter set
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2/conv_halo_r02 python tools/ncu_shapes.py stats:64,128,128,128,128 dgrad:64,128,128,128,128 wgrad:64,128,128,128,128 > gpurun_out/r2/ncu_halo.log 2>&1
python tools/ncu_digest.py gpurun_out/r2/conv_halo_r02.ncu-rep "fwd+stats 128x128 128->128 B=64 (single CTA; the bench's roofline kernel),dgrad 128x128 128->128 (cta_group::2 pairs),wgrad 128x128 128->128" --source 12 > gpurun_out/r2/conv_halo_r02.txt 2>&1
ls -la gpurun_out/r2/conv_halo_r02.ncu-rep; head -32 gpurun_out/r2/conv_halo_r02.txt
# (d) final bench lines
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_final.json 2> gpurun_out/r2/bench_final.err; cat gpurun_out/r2/bench_final.json; tail -2 gpurun_out/r2/bench_final.err
du -sh gpurun_out/r2; ls -la gpurun_out/r2 | head -50
