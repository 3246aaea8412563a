#!/bin/bash
# A/B of library variants on ONE box: tools/ab_variants.sh lib_v0 lib_v1 ...  (each hijiki_b200/<dir>/libhijiki_b200.so)
# Two rounds, interleaved, so a drifting box shows up as disagreement between the rounds.
B="--steps 4 --warmup 3 --no-cpu-baseline --no-denoiser --no-e2e --no-extras"
for round in 1 2; do
  for v in "$@"; do
    HIJIKI_B200_LIB=$PWD/hijiki_b200/$v/libhijiki_b200.so python bench.py $B ${AB_ARGS} > gpurun_out/ab_$v.$round.json 2>/dev/null
    python - <<PY
import json
j=json.loads(open("gpurun_out/ab_$v.$round.json").read().strip().splitlines()[-1])
print("$v round $round", round(j["value"]), "Mrays/s", round(j["ms_per_step"],2), "ms", {k:round(x,2) for k,x in j["kernel_ms_per_step"].items() if x}, "exact", round(j["exact_ties"]["value"]))
PY
  done
done
