#!/bin/bash
# round 2, second 8-GPU call (final build): config 2 at N=8 (weak), config 4 at N=8 and N=1 (strong-scaling end points), config 5 at N=8
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name nproc port args...
  local name=$1 np=$2 port=$3; shift 3
  if [ "$np" = 1 ]; then timeout 420 python bench.py "$@" > $O/$name.json 2> $O/$name.err
  else timeout 420 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" > $O/$name.json 2> $O/$name.err; fi
  echo "$name rc=$? $(date +%T)"
}
run r02_y_c2_n8 8 29531 --steps 5 --warmup 3
run r02_y_c4_n8 8 29532 --workload c4 --steps 2
run r02_y_c5_n8 8 29533 --workload c5 --steps 2
run r02_y_c4_n1 1 0 --workload c4 --steps 2 --no-cpu-baseline
for f in $O/r02_y_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "unavailable")}, "e2e", d.get("e2e", {}).get("value"))
    print(d.get("result", {}).get("cross_n_check"))
    print(d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
except Exception as e:
    print("unreadable", e)
PY
done
for f in $O/r02_y_*.err; do echo "== $f"; grep -v "NCCL INFO\|^$\|OMP_NUM_THREADS\|^\*\*\*" $f | tail -3; done
