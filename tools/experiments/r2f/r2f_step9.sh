#!/bin/bash
O=gpurun_out
for l in 0 1; do
  BPP_DECOMPRESS_LATE=$l BPP_ACP_TRACE=1 timeout 300 python tools/prof_round.py 52 reference-fixed 4096 16 2>&1 | grep "acp trace. verify" | tail -1 | cut -c1-300
  BPP_DECOMPRESS_LATE=$l timeout 300 python bench.py --workload shuffle --no-cpu --no-extra --steps 12 --warmup 3 > $O/r2f_late$l.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("$O/r2f_late$l.json"))
print("late $l: shuffle single", round(d["single_stream"]["value"]), "lanes", round(d["multi_lane"]["value"]), "e2e", round(d["e2e"]["value"]), "| fixed single", round(d["fixed"]["single_stream"]["value"]), "lanes", round(d["fixed"]["multi_lane"]["value"]), "e2e", round(d["fixed"]["e2e"]["value"]))
PY
done
