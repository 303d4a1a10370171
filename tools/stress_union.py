import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import cuda_selection_criteria_b200 as S
from cuda_selection_criteria_b200 import synth
from cuda_selection_criteria_b200.selection import AUX_SMH
plan = synth.make_plan(2500, 21)
regs = synth.hll(plan, 14); aux = synth.smh(plan, 128)
nctx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctxs = [S.Selection(0) for _ in range(nctx)]
for c in ctxs: c.load(regs, aux, AUX_SMH)
ref = None
for it in range(60):
    for ci, c in enumerate(ctxs):
        for tau in (0.85, 0.8):
            try:
                r = c.run(tau=np.float32(tau), criterion="smh_a")
            except Exception as e:
                print("FAIL it", it, "ctx", ci, "tau", tau, e); sys.exit(1)
            key = (tau,)
            if ref is None: ref = {}
            if key not in ref: ref[key] = (r.i.copy(), r.k.copy(), r.jaccard.copy(), r.stats["pairs_aux"])
            else:
                assert np.array_equal(ref[key][0], r.i) and np.array_equal(ref[key][2], r.jaccard), ("MISMATCH", it, ci, tau)
print("OK", nctx, {k: (v[0].size, v[3]) for k, v in ref.items()})
