#!/bin/bash
# compute-sanitizer passes over a small selection (the smoke case: smh_a and hll_a at n=512 against the oracle) and
# over the union-form A/B worker at reduced size: memcheck (out-of-bounds / misaligned), racecheck (shared-memory
# hazards of the staging rings and histogram columns), synccheck (barriers, __syncwarp masks), initcheck.
# Needs a GPU:   gpurun --timeout 900 -- 'bash tools/sanitize.sh 800'      -> gpurun_out/sanitize_*.log
# Each tool gets min(its limit, what is left of BUDGET seconds); sanitizer runs are 10-100x slower than plain ones.
BUDGET=${1:-800}
OUT=gpurun_out
mkdir -p $OUT
SMOKE='import __graft_entry__ as g; g.smoke()'
for tool in memcheck racecheck synccheck initcheck; do
    left=$(( BUDGET - SECONDS ))
    [ "$left" -lt 30 ] && { echo "$tool: skipped (only ${left}s left)"; continue; }
    lim=$(( left < 240 ? left : 240 ))
    for form in planes subsets split; do
        SELB200_UNION=$form SELB200_HLLFILTER=$form timeout -k 5 $lim compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 9 \
            python -c "$SMOKE" > $OUT/sanitize_${tool}_${form}.log 2>&1
        echo "$tool ($form): rc=$? $(grep -c 'ERROR SUMMARY' $OUT/sanitize_${tool}_${form}.log) summary line(s): $(grep 'ERROR SUMMARY' $OUT/sanitize_${tool}_${form}.log | tail -1)"
    done
done
