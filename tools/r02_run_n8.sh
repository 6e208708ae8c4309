#!/bin/bash
# round 2, 8-GPU call: BASELINE configs 4 and 5 at full size (strong scaling 8/4/2/1 of ONE 3600-frame 4K clip; the 64 x 600
# clip farm on 8 and 4 ranks) and config 2 at N=8 (end-to-end scaling after the one-upload yuv path and CPU binding)
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02_e_topo_n8.txt 2>&1
lscpu | grep -i -E "model name|socket|numa|^CPU\(s\)" >> $O/r02_e_topo_n8.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name nproc port args...
  local name=$1 np=$2 port=$3; shift 3
  if [ "$np" = 1 ]; then timeout 420 python bench.py "$@" > $O/$name.json 2> $O/$name.err
  else timeout 420 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" > $O/$name.json 2> $O/$name.err; fi
  echo "$name rc=$? $(date +%T)"
}
run r02_e_c2_n8 8 29521 --steps 5 --warmup 3
run r02_e_c4_n8 8 29522 --workload c4 --steps 2
run r02_e_c5_n8 8 29523 --workload c5 --steps 2
run r02_e_c4_n4 4 29524 --workload c4 --steps 2
run r02_e_c4_n2 2 29525 --workload c4 --steps 2
run r02_e_c5_n4 4 29526 --workload c5 --steps 2
run r02_e_c4_n1 1 0 --workload c4 --steps 2 --no-cpu-baseline
for f in $O/r02_e_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "unavailable")}, "e2e", d.get("e2e", {}).get("value"))
    print(d.get("result", {}).get("cross_n_check"))
    print(d["config"].get("cpu_binding"), d.get("clocks", {}).get("reasons"))
except Exception as e:
    print("unreadable", e)
PY
done
for f in $O/r02_e_*.err; do echo "== $f"; grep -v "NCCL INFO\|^$\|OMP_NUM_THREADS\|^\*\*\*" $f | tail -3; done
