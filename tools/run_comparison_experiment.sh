#!/usr/bin/env bash
# CPU-vs-GPU similarity comparison, the experiment of the reference's run_comparison_experiment.sh:1-55 with its
# parameters as environment variables: runs a "CPU" selection binary (flags -l -t -h -a -c smh_a) and a "GPU" one
# (-l -b -h -a), keys every output line by nameA_nameB, joins the two outputs and writes
#   cfg,card1,card2,sim_cpu,sim_gpu,diff          (diff below EPS printed as 0)
# Defaults compare this repo's two drivers; CPU_BINARY=oracle/_ref/selection puts the unmodified reference on the
# CPU side.  Lines only one side printed are reported on stderr (the reference's `join` drops them silently).
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
LISTA="${LISTA:-test_influeza_filelist.txt}"
THRESHOLD="${THRESHOLD:-0.9}"
REPS="${REPS:-1}"
THREADS="${THREADS:-8}"
BLOCK_SIZE="${BLOCK_SIZE:-128}"
MH_SIZE_ARR=(${MH_SIZE_ARR:-512})
EPS="${EPS:-1e-6}"
CPU_BINARY="${CPU_BINARY:-$ROOT/cuda_selection_criteria_b200/bin/selection}"
GPU_BINARY="${GPU_BINARY:-$ROOT/cuda_selection_criteria_b200/bin/selection_cuda}"
OUT="${OUT:-comparacion_cpu_gpu.csv}"

echo "cfg,card1,card2,sim_cpu,sim_gpu,diff" > "$OUT"
keyed () {   # run a binary, prefix every line with its join key, sort by it
    local tmp; tmp=$(mktemp)
    "$@" | awk '{print $1"_"$2, $0}' | LC_ALL=C sort -k1,1 > "$tmp"
    echo "$tmp"
}
for m in "${MH_SIZE_ARR[@]}"; do
    for r in $(seq 1 "$REPS"); do
        cpu_out=$(keyed "$CPU_BINARY" -l "$LISTA" -t "$THREADS" -h "$THRESHOLD" -a "$m" -c smh_a)
        gpu_out=$(keyed "$GPU_BINARY" -l "$LISTA" -b "$BLOCK_SIZE" -h "$THRESHOLD" -a "$m")
        LC_ALL=C join -1 1 -2 1 "$cpu_out" "$gpu_out" |
            awk -v cfg="t${THREADS}_b${BLOCK_SIZE}_m${m}_r${r}" -v eps="$EPS" '{
                c = $4; g = $7; d = (c > g ? c - g : g - c); if (d < eps) d = 0;
                print cfg "," $2 "," $3 "," c "," g "," d }' >> "$OUT"
        only_cpu=$(LC_ALL=C join -v 1 "$cpu_out" "$gpu_out" | wc -l)
        only_gpu=$(LC_ALL=C join -v 2 "$cpu_out" "$gpu_out" | wc -l)
        echo "m=$m rep=$r: pairs only on the CPU side: $only_cpu, only on the GPU side: $only_gpu" >&2
        rm -f "$cpu_out" "$gpu_out"
    done
done
echo "Comparison written: $(($(wc -l < "$OUT") - 1)) joined pairs in '$OUT'"
