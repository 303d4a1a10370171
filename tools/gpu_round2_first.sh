#!/bin/bash
# First GPU call of round 2: measures everything that was written in round 1 after the GPU budget ran out, inside a
# hard time budget, most valuable first.  One GPU.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh 1400'      -> gpurun_out/r2a_*
# 1. the whole -m gpu suite with the gated tests enabled (subset union form, packed upload)
# 2. bench.py, kernel-resident part: default plane kernel, SELB200_UNION=subsets, then the full default line
# 3. bench.py e2e: default upload against SELB200_H2D=planes; C5 hll_a with the one-hot and the subset plane filter
# 4. ncu launch list of the faster union form + one --set full capture of its union kernel (after 2 has exited 0)
# 5. compute-sanitizer passes (tools/sanitize.sh) with what is left
BUDGET=${1:-1400}
TAG=${2:-r2a}
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/${TAG}_steps.log
: > $LOG
step() {   # step NAME LIMIT cmd...
    local name=$1 lim=$2; shift 2
    local l=$(( BUDGET - SECONDS ))
    if [ "$l" -lt 10 ]; then echo "$name: skipped (only ${l}s left)" >> $LOG; return 99; fi
    [ "$lim" -gt "$l" ] && lim=$l
    local t0=$SECONDS
    timeout -k 5 "$lim" "$@"
    local rc=$?
    echo "$name: rc=$rc in $(( SECONDS - t0 ))s (limit ${lim}s)" >> $LOG
    return $rc
}
KO="python bench.py --no-cpu-baseline --no-e2e"
SELB200_TEST_SUBSETS=1 SELB200_TEST_H2D=1 step pytest 420 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
step bench_planes 60 $KO > $OUT/${TAG}_bench_planes.json 2> $OUT/${TAG}_bench_planes.err
SELB200_UNION=subsets step bench_subsets 60 $KO > $OUT/${TAG}_bench_subsets.json 2> $OUT/${TAG}_bench_subsets.err
step bench_e2e_bytes 120 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_e2e_bytes.json 2> /dev/null
SELB200_H2D=planes step bench_e2e_planes 120 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_e2e_planes.json 2> /dev/null
python - <<PY | tee $OUT/${TAG}_summary.txt
import json
def line(f):
    try:
        return json.loads(open("$OUT/${TAG}_bench_%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e)
for f in ("planes", "subsets"):
    d = line(f)
    if d:
        r = d["roofline"]
        print(f, "ms/step", round(d["ms_per_step"], 3), "union", round(r["kernels_ms"]["union"], 3), "run", round(r["kernels_ms"]["run_total"], 3),
              "pairs_aux", d["config"]["pairs_aux_rank0"], "pairs_out", d["config"]["pairs_out"])
for f in ("e2e_bytes", "e2e_planes"):
    d = line(f)
    if d:
        print(f, "e2e ms/step", round(d["e2e"]["ms_per_step"], 2), d["e2e"].get("rank0_phases_ms"), "h2d bytes", d["e2e"]["h2d_bytes_per_step"])
PY
# geometry variants of the plane kernels, when built beforehand (make -C cuda_selection_criteria_b200/csrc variants)
for lib in cuda_selection_criteria_b200/csrc/variants/libselb200_*.so; do
    [ -f "$lib" ] || continue
    tag=$(basename $lib .so); tag=${tag#libselb200_}
    for form in planes subsets; do
        SELB200_LIB=$PWD/$lib SELB200_UNION=$form step bench_${tag}_$form 45 $KO > $OUT/${TAG}_bench_${tag}_$form.json 2> /dev/null
        python - <<PY | tee -a $OUT/${TAG}_summary.txt
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_${tag}_$form.json").read().strip().splitlines()[-1])
    print("$tag $form", "ms/step", round(d["ms_per_step"], 3), "union", round(d["roofline"]["kernels_ms"]["union"], 3))
except Exception as e:
    print("$tag $form unreadable:", e)
PY
    done
done
# C5 (n=50k, hll_a, p_aux=10): plane hll filter, one-hot against subset counting
C5="python bench.py --steps 5 --warmup 3 --n 50000 --seed 1003 --criterion hll_a --aux-bytes 1024 --no-cpu-baseline --no-e2e"
step bench_c5_planes 90 $C5 > $OUT/${TAG}_bench_c5_planes.json 2> /dev/null
SELB200_HLLFILTER=subsets step bench_c5_subsets 90 $C5 > $OUT/${TAG}_bench_c5_subsets.json 2> /dev/null
python - <<PY | tee -a $OUT/${TAG}_summary.txt
import json
for f in ("c5_planes", "c5_subsets"):
    try:
        d = json.loads(open("$OUT/${TAG}_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 3), "filter", round(d["roofline"]["kernels_ms"]["filter"], 3), "pairs_aux", d["config"]["pairs_aux_rank0"], "pairs_out", d["config"]["pairs_out"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
# ncu only for the form that won, and only if its plain run exited 0
FORM=planes
python - <<PY && FORM=subsets
import json, sys
a = json.loads(open("$OUT/${TAG}_bench_planes.json").read().strip().splitlines()[-1])["roofline"]["kernels_ms"]["union"]
b = json.loads(open("$OUT/${TAG}_bench_subsets.json").read().strip().splitlines()[-1])["roofline"]["kernels_ms"]["union"]
sys.exit(0 if b < a else 1)
PY
echo "ncu on SELB200_UNION=$FORM" >> $LOG
SELB200_UNION=$FORM step ncu_launches 180 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_ncu_launches_$FORM.csv $KO --steps 2 --warmup 3 > $OUT/${TAG}_ncu_launches.log 2>&1
SELB200_UNION=$FORM step ncu_full 240 ncu --set full --clock-control none --import-source on -k regex:k_pair_hist_planes -c 1 \
    -o $OUT/${TAG}_union_$FORM $KO --steps 1 --warmup 1 > $OUT/${TAG}_ncu_full.log 2>&1
left=$(( BUDGET - SECONDS - 10 ))
[ "$left" -gt 60 ] && bash tools/sanitize.sh $left > $OUT/${TAG}_sanitize.txt 2>&1
cat $LOG
tail -3 $OUT/${TAG}_pytest.log
