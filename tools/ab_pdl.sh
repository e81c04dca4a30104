#!/bin/bash
# Same-box A/B of programmatic dependent launch (RFV_FLAG_NO_PDL = 16777216 as the control): unprofiled forward time of the
# per-layer tool at batch 64 (launch-bound) and at micro-batch 512.
for i in 1 2; do
  for mb in 64 512; do
    python tools/profile_layers.py --mb $mb --reps 20 2>&1 | grep -E "^micro" | sed "s/^/pdl:    /"
    python tools/profile_layers.py --mb $mb --reps 20 --flags 16777216 2>&1 | grep -E "^micro" | sed "s/^/serial: /"
  done
done
