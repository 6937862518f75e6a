#!/bin/bash
# A/B of the single-precision pre-cull on the bench workload (same box, back to back)
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/b_cull.json
UPC_B200_NO_FP32_CULL=1 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/b_nocull.json
python - <<'PY'
import json
for f in ("b_cull", "b_nocull"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "sim kernel ms %.3f" % d["roofline"]["kernel_ms"])
PY
