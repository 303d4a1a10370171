#!/bin/bash
# Short A/B of the union-pass forms on one B200 (inside a hard time budget): the A/B and split-specific GPU tests,
# then bench.py (kernel-resident part only) with SELB200_UNION=split, with the 20-CTAs/SM build of the split kernel
# (csrc/variants/libselb200_c20.so, built by hand with -DSPLIT_MIN_CTAS=20) and with the default plane kernel.
# usage: tools/gpu_ab_split.sh [BUDGET_SECONDS] [TAG]
BUDGET=${1:-55}
TAG=${2:-s81}
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/${TAG}_steps.log
: > $LOG
step() {   # step NAME LIMIT cmd...
    local name=$1 lim=$2; shift 2
    local l=$(( BUDGET - SECONDS ))
    if [ "$l" -lt 6 ]; then echo "$name: skipped (only ${l}s left)" >> $LOG; return 99; fi
    [ "$lim" -gt "$l" ] && lim=$l
    local t0=$SECONDS
    timeout -k 2 "$lim" "$@"
    local rc=$?
    echo "$name: rc=$rc in $(( SECONDS - t0 ))s (limit ${lim}s)" >> $LOG
    return $rc
}
B="python bench.py --no-cpu-baseline --no-e2e"
step pytest 34 python -m pytest tests/test_gpu_ab.py tests/test_gpu_parity.py::test_split_union_mixed_bases_and_long_lists -x -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
SELB200_UNION=split step bench_split 12 $B > $OUT/${TAG}_bench_split.json 2> /dev/null
C20=$PWD/cuda_selection_criteria_b200/csrc/variants/libselb200_c20.so     # optional: nvcc ... -DSPLIT_MIN_CTAS=20 build
[ -f $C20 ] && SELB200_UNION=split SELB200_LIB=$C20 step bench_split_c20 12 $B > $OUT/${TAG}_bench_split_c20.json 2> /dev/null
step bench_planes 12 $B > $OUT/${TAG}_bench_planes.json 2> /dev/null
cat $LOG
tail -2 $OUT/${TAG}_pytest.log
python - <<PY
import json
for f in ("split", "split_c20", "planes"):
    try:
        d = json.loads(open("$OUT/${TAG}_bench_%s.json" % f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, d["ms_per_step"], d["config"]["pairs_aux_rank0"], d["config"]["pairs_out"], r["kernel"], r["kernels_ms"]["union"], r["kernels_ms"]["run_total"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
