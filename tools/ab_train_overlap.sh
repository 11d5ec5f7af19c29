#!/bin/bash
# A/B of the overlapped gradient all-reduce of the training step (TQ_TRAIN_OVERLAP) at N GPUs: bench.py's cfg4 leg only.
N=${1:-2}

for ov in 1 0 1 0; do
  TQ_TRAIN_OVERLAP=$ov python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29560 + ov)) \
      bench.py --gpus $N --steps 3 --warmup 3 --configs cfg4 2>/dev/null | tail -1 > /tmp/ab_$ov.json
  python - "$ov" <<'PY'
import json, sys
d = json.loads(open(f"/tmp/ab_{sys.argv[1]}.json").read())
c = d["configs"]["cfg4"]
print("TQ_TRAIN_OVERLAP", sys.argv[1], "ms_per_step", round(c["ms_per_step"], 3), "samples/s", round(c["value"], 1), flush=True)
PY
done
