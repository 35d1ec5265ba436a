#!/bin/bash
# A/B a library option on the C4 slab geometry (one GPU, no communication)
opt=$1
for f in 1 0 1 0; do
  timeout 200 python bench.py --workload c4 --shape 160,192,192 --steps 3 --warmup 3 --opt $opt=$f 2>/dev/null | tail -1 > /tmp/c4ab.json
  python - "$f" <<'P'
import json, sys
d = json.load(open("/tmp/c4ab.json")); b = d["kernel_breakdown_ms_per_step_rank0"]
print("opt =", sys.argv[1], "ms/step", round(d["ms_per_step"], 2), {k: round(v["ms"], 2) for k, v in b.items()}, d["clocks"]["sm_mhz"])
P
done
