#!/bin/bash
# after a kernel change: GPU test suite, the default bench line, the ncu launch list with DRAM bytes (-> profiles/ncu_traffic.json)
set -u
O=gpurun_out
T=${1:-j}
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_gpu.log
tail -4 $O/r02_${T}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r02_${T}_bench_c2.json 2> $O/r02_${T}_bench_c2.err; echo "c2 rc=$?"
python - "$O/r02_${T}_bench_c2.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d['roofline']
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'roof', r['kernel'], round(r['frac'], 3), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['result']['scene_complexity'][0])
for k, v in list(r['kernels'].items())[:12]: print('   ', k, v['ms'], v['launches'], v['frac_of_hbm_peak'])
PY
VQA_PROF_REPORT=$O/r02_${T}_step_report.json timeout 300 python tools/profile_step.py 24 > $O/prof_plain.log 2>&1 || tail -5 $O/prof_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file $O/r02_${T}_ncu_launches_step24.csv python tools/profile_step.py 24 > $O/r02_${T}_ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/traffic_from_csv.py $O/r02_${T}_ncu_launches_step24.csv $O/r02_${T}_step_report.json r02_${T}_ncu_launches_step24.csv | tee $O/r02_${T}_traffic.txt | head -12
cp profiles/ncu_traffic.json $O/r02_${T}_ncu_traffic.json
