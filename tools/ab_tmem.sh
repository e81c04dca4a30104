#!/bin/bash
# Same-box A/B of the tensor-memory-resident weight blocks of conv_wa (RFV_WA_TMEM = 0 off / 1 default / 2 every layer it fits):
# per-layer profile at micro-batch 512, two alternating rounds; full tables of the last round are kept.
for i in 1 2; do
  for m in 0 1 2; do
    RFV_WA_TMEM=$m timeout 120 python tools/profile_layers.py --mb 512 > /tmp/tm_$m.log 2>&1
    grep -E "^micro|^attention=" /tmp/tm_$m.log | cut -c1-260 | sed "s/^/tmem=$m: /"
  done
done
for m in 0 1 2; do echo "==== RFV_WA_TMEM=$m"; grep -E "^conv_wa" /tmp/tm_$m.log | sort -k2; done
