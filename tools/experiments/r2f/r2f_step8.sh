#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2f_pytest8.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2f_pytest8.log
timeout 300 python bench.py --workload shuffle --no-cpu --no-extra --no-fixed --steps 12 --warmup 3 > $O/r2f_wit.json 2>/dev/null
python - <<PY
import json
d=json.load(open("$O/r2f_wit.json"))
print("shuffle single", round(d["single_stream"]["value"]), "lanes", round(d["multi_lane"]["value"]), "e2e", round(d["e2e"]["value"]), "serial e2e", round(d["e2e"]["serial"]["value"]))
PY
