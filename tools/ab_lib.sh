#!/bin/bash
# Same-box A/B of two builds of the library: default vs RFV_LIB=<path>; alternates 3 rounds of the per-layer profile.
ALT=${1:-rectified_flow_vision_b200/librfv_b200_ab.so}
for i in 1 2 3; do
  python tools/profile_layers.py --mb 512 2>&1 | grep -E "^micro|^attention" | sed "s/^/default: /"
  RFV_LIB=$PWD/$ALT python tools/profile_layers.py --mb 512 2>&1 | grep -E "^micro|^attention" | sed "s/^/alt:     /"
done
