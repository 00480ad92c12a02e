#!/bin/bash
# exposed-exchange sweep at N ranks (default 2): all-reduce backend, CTAs and bucket size; usage: tools/n2_sweep.sh [N] VAR=VALUE ...
N=${1:-2}; shift
run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],3))"; }
for cfg in "$@"; do run $cfg; done
