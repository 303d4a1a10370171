#!/usr/bin/env python
"""Per-kernel device times of one shard of the C4 workload on ONE GPU (what each rank of an N-GPU job runs).

    python tools/shard_probe.py --shards 8 [--n 100000] [--criterion smh_a]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_selection_criteria_b200 as S  # noqa: E402
from cuda_selection_criteria_b200 import synth  # noqa: E402
from cuda_selection_criteria_b200.selection import AUX_HLL, AUX_NONE, AUX_SMH  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--shards", type=int, default=8)
ap.add_argument("--criterion", default="smh_a")
ap.add_argument("--aux-bytes", type=int, default=1024)
ap.add_argument("--tau", type=float, default=0.9)
ap.add_argument("--seed", type=int, default=1002)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()

plan = synth.make_plan(a.n, a.seed)
regs = synth.hll(plan, 14, device=0)
kind = {"cb": AUX_NONE, "smh_a": AUX_SMH, "hll_a": AUX_HLL, "hll_an": AUX_HLL}[a.criterion]
aux = None
if kind == AUX_SMH:
    aux = synth.smh(plan, a.aux_bytes // 8, device=0)
elif kind == AUX_HLL:
    aux = synth.hll(plan, a.aux_bytes.bit_length() - 1, synth.TAG_AUX_HLL, device=0)
sel = S.Selection(0)
sel.load(regs, aux, kind)
keys = ("ms_bounds", "ms_filter", "ms_verify", "ms_union", "ms_estimate", "ms_sort", "ms_total")
for shards in sorted({1, a.shards}):
    for shard in sorted({0, shards - 1}):
        acc = []
        for _ in range(a.reps):
            st = sel.run(tau=np.float32(a.tau), criterion=a.criterion, shard=shard, n_shards=shards, fetch=False).stats
            acc.append(st)
        best = {k: min(s[k] for s in acc[1:]) for k in keys}
        print(f"shard {shard}/{shards}: pairs_cb_shard={st['pairs_cb_shard']} pairs_aux={st['pairs_aux']} out={st['pairs_out']} "
              + " ".join(f"{k[3:]}={v:.3f}" for k, v in best.items()), flush=True)
