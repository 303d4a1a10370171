set -x
B="python bench.py --steps 5 --warmup 3"
$B --n 10000 --seed 1001 --criterion cb --tau 0.9 --cpu-sample 10000 > gpurun_out/cfg_C2.json 2>/dev/null
for t in 0.70 0.75 0.80 0.85 0.90 0.95; do $B --n 10000 --seed 1001 --criterion smh_a --tau $t --no-cpu-baseline --no-e2e > gpurun_out/cfg_C3_$t.json 2>/dev/null; done
$B --n 50000 --seed 1003 --criterion hll_a --aux-bytes 1024 --cpu-sample 12000 > gpurun_out/cfg_C5_hlla_p10.json 2>/dev/null
$B --n 50000 --seed 1003 --criterion hll_an --aux-bytes 1024 --cpu-sample 12000 > gpurun_out/cfg_C5_hllan_p10.json 2>/dev/null
$B --n 50000 --seed 1003 --criterion hll_a --aux-bytes 256 --no-cpu-baseline > gpurun_out/cfg_C5_hlla_p8.json 2>/dev/null
for f in gpurun_out/cfg_*.json; do python - "$f" <<'PY'
import sys,json
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
cb=d.get("cpu_baseline") or {}
print(sys.argv[1].split("/")[-1], "ms/step", round(d["ms_per_step"],3), "pairs/s", f'{d["value"]:.3e}', "e2e_ms", (d.get("e2e") or {}).get("ms_per_step"), "P_cb", d["config"]["pairs_cb"], "P_aux", d["config"]["pairs_aux_rank0"], "P_out", d["config"]["pairs_out"], "kernels", {k: round(v,3) for k,v in d["roofline"]["kernels_ms"].items()}, "cpu", f'{cb.get("value",0):.3e}', cb.get("cores"))
PY
done
