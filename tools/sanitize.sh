#!/bin/bash
# compute-sanitizer over the small-shape driver, one report per tool and kernel family:
#   bash tools/sanitize.sh [out_dir]      (on the GPU box; summaries go to profiles/ by hand)
# memcheck = out-of-bounds / misaligned accesses, racecheck = shared-memory hazards,
# synccheck = invalid barrier usage, initcheck = reads of uninitialised global memory.
OUT="${1:-gpurun_out/sanitizer}"
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # tool, family, timeout
  local log="$OUT/$1_$2.log"
  timeout "$3" "$CS" --tool "$1" --print-limit 20 --error-exitcode 77 \
      python tools/sanitize_driver.py "$2" > "$log" 2>&1
  local rc=$?
  echo "$1 $2: rc=$rc  $(grep -c '^ok ' "$log") checks ok  |  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tail -1)"
}
for fam in screen tf32x3 simt rescale analysis; do run memcheck "$fam" "${KB2_SAN_TIMEOUT:-420}"; done
for fam in screen tf32x3 rescale; do run synccheck "$fam" "${KB2_SAN_TIMEOUT:-420}"; done
for fam in screen tf32x3 rescale; do run racecheck "$fam" "${KB2_SAN_TIMEOUT:-420}"; done
