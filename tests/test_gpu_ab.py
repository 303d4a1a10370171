"""-m gpu: the forms of the union pass (bit planes with subset counting = default, with one-hot counting, bytes) and the
forms of the hll filter (two passes on bit planes = default, one pass on bit planes, bytes) against each other, at sizes
the CPU oracle would take minutes for.  All forms compute the same integer histograms and the same fp64 decision, so
pair lists, Jaccard bits and stage counts must be identical (SELB200_UNION=planes|bytes is read when a context is
created, SELB200_HLLFILTER=onepass|bytes at library load)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import hashlib, json, os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import cuda_selection_criteria_b200 as S
from cuda_selection_criteria_b200 import synth
from cuda_selection_criteria_b200.selection import AUX_HLL, AUX_SMH
out = {}
plan = synth.make_plan(20000, 1002)
regs = synth.hll(plan, 14, device=0)
for crit, aux, kind, tau in (("smh_a", synth.smh(plan, 128, device=0), AUX_SMH, 0.9),
                             ("smh_a", synth.smh(plan, 128, device=0), AUX_SMH, 0.75),
                             ("hll_a", synth.hll(plan, 10, synth.TAG_AUX_HLL, device=0), AUX_HLL, 0.9),
                             ("hll_an", synth.hll(plan, 8, synth.TAG_AUX_HLL, device=0), AUX_HLL, 0.85)):
    with S.Selection(0) as sel:
        sel.load(regs, aux, kind)
        r = sel.run(tau=np.float32(tau), criterion=crit)
    h = hashlib.sha256(r.i.tobytes() + r.k.tobytes() + r.jaccard.tobytes()).hexdigest()
    out[f"{crit}@{tau}"] = [h, int(r.i.size), int(r.stats["pairs_cb"]), int(r.stats["pairs_aux"]), int(r.stats["pairs_near"])]
# 1500 copies of one genome: every band of theirs is one bucket of 1500 (1.1 M items per band, 18 M in all: the join's item
# list overflows its first capacity and the pass is redone), 1.1 M pairs with J = 1 (the output list grows too)
plan = synth.make_plan(4000, 7)
regs_h, aux_h = synth.hll(plan, 14).copy(), synth.smh(plan, 128).copy()
regs_h[2000:3500] = regs_h[1999]
aux_h[2000:3500] = aux_h[1999]
with S.Selection(0) as sel:
    sel.load(regs_h, aux_h, AUX_SMH)
    r = sel.run(tau=np.float32(0.9), criterion="smh_a")
h = hashlib.sha256(r.i.tobytes() + r.k.tobytes() + r.jaccard.tobytes()).hexdigest()
out["duplicates"] = [h, int(r.i.size), int(r.stats["pairs_cb"]), int(r.stats["pairs_aux"]), int(r.stats["pairs_near"])]
print("RESULT " + json.dumps(out))
"""


def _run(env_extra, tmp_path):
    script = tmp_path / "ab.py"
    script.write_text(_WORKER)
    env = dict(os.environ)
    env.pop("SELB200_UNION", None)
    env.pop("SELB200_HLLFILTER", None)
    env.pop("SELB200_SMHFILTER", None)
    env.update(env_extra)
    r = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [x for x in r.stdout.splitlines() if x.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_plane_kernels_equal_byte_kernels(gpu, tmp_path):
    default = _run({}, tmp_path)
    onehot = _run({"SELB200_UNION": "planes"}, tmp_path)
    by = _run({"SELB200_UNION": "bytes", "SELB200_HLLFILTER": "bytes"}, tmp_path)
    assert default == by
    assert onehot == by
    onepass = _run({"SELB200_HLLFILTER": "onepass"}, tmp_path)
    assert onepass == by
    tiles = _run({"SELB200_SMHFILTER": "tiles"}, tmp_path)        # all-pairs tile filter + verify instead of the equality join
    assert tiles == by
    # the subset union kernel without its per-step group limit (per-eighth maxima), and the wide pairs on the side stream
    notops = _run({"SELB200_UNION_TOPS": "0"}, tmp_path)
    assert notops == by
    side = _run({"SELB200_WIDE": "side"}, tmp_path)
    assert side == by
    assert all(v[1] > 1000 for v in default.values())          # thousands of emitted pairs in every case
    assert default["duplicates"][1] > 1500 * 1499 // 2         # every pair of the 1501 identical genomes, and the rest
