#!/bin/bash
# round 2 consolidation run (1 GPU): full GPU test suite, the bench lines of c2 / c1 / c3 for both arms, the ncu launch list of one
# step with DRAM bytes per launch (-> profiles/ncu_traffic.json through tools/traffic_from_csv.py)
set -u
O=gpurun_out
T=${1:-i}
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_gpu.log
tail -6 $O/r02_${T}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r02_${T}_bench_c2.json 2> $O/r02_${T}_bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --workload c1 > $O/r02_${T}_bench_c1.json 2> $O/r02_${T}_bench_c1.err; echo "c1 rc=$?"
timeout 300 python bench.py --workload c3 > $O/r02_${T}_bench_c3.json 2> $O/r02_${T}_bench_c3.err; echo "c3 rc=$?"
timeout 400 python bench.py --impl reference --workload c1 --steps 2 --warmup 1 > $O/r02_${T}_ref_c1.json 2> $O/r02_${T}_ref_c1.err; echo "ref c1 rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_${T}_ref_c2.json 2> $O/r02_${T}_ref_c2.err; echo "ref c2 rc=$?"
for w in c2 c1 c3; do python - "$O/r02_${T}_bench_$w.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d['roofline']
print(sys.argv[1], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'],
      'roof', r['kernel'], round(r['frac'], 3), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'], 2), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
cut -c1-600 $O/r02_${T}_ref_c1.json; cut -c1-400 $O/r02_${T}_ref_c2.json
VQA_PROF_REPORT=$O/r02_${T}_step_report.json timeout 300 python tools/profile_step.py 24 > $O/prof_plain.log 2>&1 || tail -5 $O/prof_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file $O/r02_${T}_ncu_launches_step24.csv python tools/profile_step.py 24 > $O/r02_${T}_ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/traffic_from_csv.py $O/r02_${T}_ncu_launches_step24.csv $O/r02_${T}_step_report.json r02_${T}_ncu_launches_step24.csv | tee $O/r02_${T}_traffic.txt | head -14
cp profiles/ncu_traffic.json $O/r02_${T}_ncu_traffic.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02_${T}_smoke.log
