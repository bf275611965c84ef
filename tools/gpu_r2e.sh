#!/bin/bash
for v in 0 4; do for sh in "1024 1024 20" "4096 4096 10" "1024 77 20"; do timeout 120 python tools/attn_stall.py $v $sh 200; done; done
