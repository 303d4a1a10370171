#!/bin/bash
# One gpurun call that validates a union-kernel change end to end inside a hard time budget:
#   smoke -> pytest -m gpu -> bench.py (defaults) -> ncu launch list -> ncu --set full of the union kernel.
# Every step gets min(its own limit, what is left of BUDGET seconds); a failing smoke switches to diagnostics
# (SELB200_DEBUG_SYNC + compute-sanitizer on the smoke case) instead of spending the budget on hanging tests.
# usage: tools/gpu_validate_split.sh [BUDGET_SECONDS] [TAG]
BUDGET=${1:-185}
TAG=${2:-s80}
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/${TAG}_steps.log
: > $LOG
left() { echo $(( BUDGET - SECONDS )); }
step() {   # step NAME LIMIT cmd...
    local name=$1 lim=$2; shift 2
    local l=$(left)
    if [ "$l" -lt 12 ]; then echo "$name: skipped (only ${l}s left)" >> $LOG; return 99; fi
    [ "$lim" -gt "$l" ] && lim=$l
    local t0=$SECONDS
    timeout -k 3 "$lim" "$@"
    local rc=$?
    echo "$name: rc=$rc in $(( SECONDS - t0 ))s (limit ${lim}s)" >> $LOG
    return $rc
}

step smoke 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
if [ $? -ne 0 ]; then
    echo "smoke failed: diagnostics only" >> $LOG
    SELB200_DEBUG_SYNC=1 step smoke_dbg 40 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke_dbg.log 2>&1
    step sanitizer 80 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_sanitizer.log 2>&1
    SELB200_UNION=planes step smoke_planes 30 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke_planes.log 2>&1
    cat $LOG; tail -5 $OUT/${TAG}_smoke.log
    exit 1
fi
step pytest 95 python -m pytest tests -m gpu -x -q --durations=8 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
PYRC=$?
step bench 45 python bench.py > $OUT/${TAG}_bench1.json 2> $OUT/${TAG}_bench1.err
if [ $PYRC -eq 0 ]; then
    step ncu_list 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_${TAG}_list.log 2>&1
    step ncu_full 45 ncu --set full --clock-control none --import-source on -k regex:k_pair_hist_split -c 1 -f -o $OUT/prof_${TAG}_split \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_${TAG}_full.log 2>&1
fi
SELB200_UNION=planes step bench_planes 25 python bench.py --no-cpu-baseline --no-e2e > $OUT/${TAG}_bench1_planes.json 2> /dev/null
cat $LOG
tail -3 $OUT/${TAG}_pytest.log
python - <<EOF
import json
for f in ("$OUT/${TAG}_bench1.json", "$OUT/${TAG}_bench1_planes.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, d["ms_per_step"], d["value"], d["config"]["pairs_aux_rank0"], d["config"]["pairs_out"], r["kernel"], r["kernels_ms"],
              (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "unreadable:", e)
EOF
exit $PYRC
