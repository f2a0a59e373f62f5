#!/bin/bash
for e in 0 1; do DRB_LIN_PERS=$e timeout 300 python profiles/experiments/clip_cost.py 2>&1 | grep "ms per call"; done
DRB_LIN_PERS=1 DRB_NO_PDL=1 timeout 300 python profiles/experiments/clip_cost.py 2>&1 | grep "ms per call" | sed 's/^/NO_PDL /'
