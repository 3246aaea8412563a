#!/bin/bash
# A/B of library variants on ONE box: tools/ab_variants.sh lib lib_x:coop_batch_cost=140 ...
# (each hijiki_b200/<dir>/libhijiki_b200.so, optionally with HJK_OPTIONS after a colon).  AB_ARGS = extra bench.py arguments,
# AB_ROUNDS = interleaved rounds (default 2: a drifting box shows up as disagreement between the rounds).
B="--steps 4 --warmup 3 --no-cpu-baseline --no-denoiser --no-e2e --no-extras"
for round in $(seq 1 ${AB_ROUNDS:-2}); do
  for spec in "$@"; do
    v=${spec%%:*}; opts=""; [[ "$spec" == *:* ]] && opts=${spec#*:}
    HJK_OPTIONS="$opts" HIJIKI_B200_LIB=$PWD/hijiki_b200/$v/libhijiki_b200.so python bench.py $B ${AB_ARGS} > gpurun_out/ab_tmp.json 2>/dev/null
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print("$spec round $round ${AB_ARGS}:", round(j["value"]), "Mrays/s", round(j["ms_per_step"],2), "ms", {k:round(x,2) for k,x in j["kernel_ms_per_step"].items() if x}, "exact", round(j["exact_ties"]["value"]))
except Exception as e:
    print("$spec round $round: FAILED", e)
PY
  done
done
