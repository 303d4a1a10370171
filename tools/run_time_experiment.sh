#!/usr/bin/env bash
# Timing experiment, the sweep of the reference's run_time_experiment.sh:1-44 with its parameters as environment
# variables: for every (threads | block size, M, repetition) runs the CPU-flag harness (-l -t -h -m) and the CUDA-flag
# harness (-l -h -m -b) and collects their `list;phase;tau;seconds` lines into
#   impl,threads,mh_size,rep,criterio,tiempo
# Both default to this repo's binaries (bin/time_smh and bin/time_smh_cuda run the same GPU path and time it to
# completion); CPU_BINARY may point at a build of the reference's experiments/src/time_smh.cpp.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
LISTA="${LISTA:-test_influeza_filelist.txt}"
THRESHOLD="${THRESHOLD:-0.9}"
REPS="${REPS:-1}"
THREADS_ARR=(${THREADS_ARR:-8})
BLOCK_SIZE=(${BLOCK_SIZE:-256})
MH_SIZE_ARR=(${MH_SIZE_ARR:-512})
CPU_BINARY="${CPU_BINARY:-$ROOT/cuda_selection_criteria_b200/bin/time_smh}"
GPU_BINARY="${GPU_BINARY:-$ROOT/cuda_selection_criteria_b200/bin/time_smh_cuda}"
LOG="${LOG:-experimento_smh_comparativo.csv}"

echo "impl,threads,mh_size,rep,criterio,tiempo" > "$LOG"
collect () {   # impl, threads-or-block, M, rep, harness output
    local impl=$1 t=$2 m=$3 r=$4 output=$5
    for crit in build_smh smh_a CB+smh_a; do
        echo "$output" | grep -F ";$crit;" | awk -F';' -v i="$impl" -v t="$t" -v m="$m" -v r="$r" -v c="$crit" \
            '{print i","t","m","r","c","$4}' >> "$LOG"
    done
}
for T in "${THREADS_ARR[@]}"; do for M in "${MH_SIZE_ARR[@]}"; do for REP in $(seq 1 "$REPS"); do
    collect cpu "$T" "$M" "$REP" "$("$CPU_BINARY" -l "$LISTA" -t "$T" -h "$THRESHOLD" -m "$M")"
done; done; done
for B in "${BLOCK_SIZE[@]}"; do for M in "${MH_SIZE_ARR[@]}"; do for REP in $(seq 1 "$REPS"); do
    collect gpu "$B" "$M" "$REP" "$("$GPU_BINARY" -l "$LISTA" -h "$THRESHOLD" -m "$M" -b "$B")"
done; done; done
echo "Done, results in $LOG"
