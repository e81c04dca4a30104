#!/bin/bash
# Same-box A/B of two library builds on the training step (per-kernel-class split of tools/profile_train.py).
ALT=${1:-rectified_flow_vision_b200/librfv_b200_ab.so}
for i in 1 2; do
  python tools/profile_train.py --mb 256 2>&1 | grep -E "^micro_batch|^  gn_bwd|^  gn_apply" | cut -c1-120 | sed "s/^/default: /"
  RFV_LIB=$PWD/$ALT python tools/profile_train.py --mb 256 2>&1 | grep -E "^micro_batch|^  gn_bwd|^  gn_apply" | cut -c1-120 | sed "s/^/alt:     /"
done
