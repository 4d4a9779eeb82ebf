#!/bin/bash
O=gpurun_out
for l in 3 4 6; do
  BPP_LANES=$l timeout 300 python bench.py --workload shuffle --no-cpu --no-extra --steps 12 --warmup 3 > $O/r2f_lanes$l.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("$O/r2f_lanes$l.json"))
print("lanes $l: shuffle single", round(d["single_stream"]["value"]), "lanes", round(d["multi_lane"]["value"]), "e2e", round(d["e2e"]["value"]), "| fixed single", round(d["fixed"]["single_stream"]["value"]), "lanes", round(d["fixed"]["multi_lane"]["value"]), "e2e", round(d["fixed"]["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), "traffic", d["roofline"]["traffic"])
PY
done
