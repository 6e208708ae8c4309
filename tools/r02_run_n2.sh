#!/bin/bash
# round 2, 2-GPU validation call: the 2-rank NCCL test, then reduced c2 / c4 / c5 at N=2 (and c4 at N=1 on the same clip for
# the cross-N identity of the integer rows)
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02_d_topo_n2.txt 2>&1
timeout 600 python -m pytest tests/test_yuv_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "two_rank or 2rank or nccl or farneback or golden or halo" > $O/r02_d_pytest_n2.log 2>&1; echo "pytest rc=$?" >> $O/r02_d_pytest_n2.log
tail -4 $O/r02_d_pytest_n2.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_d_c2_n2.json 2> $O/r02_d_c2_n2.err; echo "c2 n2 rc=$?"
timeout 300 python bench.py --workload c4 --frames 240 --steps 2 --no-cpu-baseline > $O/r02_d_c4_240_n1.json 2> $O/r02_d_c4_240_n1.err; echo "c4 n1 rc=$?"
timeout 300 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --workload c4 --frames 240 --steps 2 > $O/r02_d_c4_240_n2.json 2> $O/r02_d_c4_240_n2.err; echo "c4 n2 rc=$?"
timeout 300 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --workload c5 --clips 6 --frames 200 --steps 2 > $O/r02_d_c5_6x200_n2.json 2> $O/r02_d_c5_6x200_n2.err; echo "c5 n2 rc=$?"
timeout 300 python bench.py --workload c5 --clips 6 --frames 200 --steps 2 > $O/r02_d_c5_6x200_n1.json 2> $O/r02_d_c5_6x200_n1.err; echo "c5 n1 rc=$?"
for f in $O/r02_d_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d.get("e2e", {}).get("value"))
    print(d.get("result", {}).get("cross_n_check"))
    print(d["config"].get("cpu_binding"))
except Exception as e:
    print("unreadable", e)
PY
done
for f in $O/r02_d_*.err; do echo "== $f"; grep -v "NCCL INFO\|^$" $f | tail -4; done
