#!/bin/bash
# A/B of environment switches on the decoder train step: tools/ab_env.sh "A=1 B=2" "A=0" ...  -> one kernel_share line per setting
for cfg in "$@"; do  # an empty string = the defaults
  out=$(env $cfg timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1)
  python - "$cfg" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2]); k = d["kernel_share"]
print("%-40s step %.3f fwd %.3f bwd %.3f outside %.3f | no-overlap step %.3f bwd %.3f" % (sys.argv[1], k["step_ms"], k["fwd_loop_ms"], k["bwd_loop_ms"], k["outside_loops_ms"], k["without_overlap"]["step_ms"], k["without_overlap"]["bwd_loop_ms"]))
PY
done
