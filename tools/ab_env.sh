#!/bin/bash
# Same-box sweep of the weights-as-A kernel's experiment switches (environment variables read at plan / launch time).
run() { echo "== $*"; env "$@" python tools/profile_layers.py --mb 512 2>&1 | grep -E "^micro|^attention=" | cut -c1-250; }
run RFV_NONE=1
run RFV_WA_PF=1
run RFV_WA_PF=2
run RFV_WA_RS=7
run RFV_WA_WST=4
run RFV_NONE=1
