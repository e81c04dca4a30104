#!/bin/bash
# A/B of the bucketed (overlapped) gradient all-reduce against one blocking call: training step of bench.py on N GPUs.
#   bash tools/ab_buckets.sh [N]
N=${1:-2}
for b in 1 0; do
  RFV_BUCKETS=$b python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + b)) \
      bench.py --gpus $N --steps 1 --warmup 3 --no-128 2>/dev/null | grep '^{' | tail -1 > /tmp/ab_$b.json
  python - "$b" <<'PY'
import json, sys
d = json.load(open(f"/tmp/ab_{sys.argv[1]}.json"))
print("RFV_BUCKETS=" + sys.argv[1], d["summary"])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 tools/check_dist.py 2>&1 | tail -6
