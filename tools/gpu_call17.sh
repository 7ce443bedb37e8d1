set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 --timeout-method thread 2>&1 | tail -6
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -4
for cfg in phiseg_7_5 probunet phiseg_7_5_256 phiseg_7_5_gn; do
timeout 150 python bench.py --steps 10 --warmup 4 --no-cpu --config $cfg > gpurun_out/r2/bench6_$cfg.json 2> gpurun_out/r2/bench6_$cfg.err
echo "rc=$? $cfg"
python -c "
import json
d=json.load(open('gpurun_out/r2/bench6_$cfg.json'))
print('$cfg ms_per_step', d['ms_per_step'], 'img/s', d['value'], 'e2e', d['e2e']['ms_per_step'], 'sampling', d['sampling']['value'], 'kernel', d['roofline']['us_per_launch'], d['roofline']['frac'], 'step frac', d['roofline_step']['frac'])
"
done
