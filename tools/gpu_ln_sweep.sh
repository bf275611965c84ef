#!/bin/bash
# LayerNorm warps-per-CTA sweep (SDXL_B200_LN_WARPS): step time and eager LayerNorm time per setting
for w in 4 8 16; do
  SDXL_B200_LN_WARPS=$w timeout 120 python bench.py --no-cpu-baseline --steps 12 --warmup 3 2>/dev/null > /tmp/b_$w.json
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
d = json.load(open(f"/tmp/b_{w}.json"))
print("warps", w, round(d["ms_per_step"], 3), d["roofline"]["by_kernel_ms"]["layer_norm"])
PY
done
