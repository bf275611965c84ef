#!/bin/bash
# per-K-block time of the GEMM main loop vs pipeline depth (and with MMAs skipped): is it latency x bytes-in-flight bound?
for nst in 3 4 5 6 7; do
  echo "== NST=$nst"; SDXL_B200_DBG_NST=$nst python tools/igemm_timeline.py 2>&1 | grep -E "^2048 11520|^2048 5120|^2048 1280 1280"
done
for nst in 4 7; do
  echo "== NST=$nst loads only (MMAs skipped)"; SDXL_B200_DBG_MODE=2 SDXL_B200_DBG_NST=$nst python tools/igemm_timeline.py 2>&1 | grep -E "^2048 11520|^2048 5120"
done
