"""Synthetic inputs of the reference-binary goldens (shared by make_golden.py and the tests)."""
from __future__ import annotations

import hashlib

import numpy as np

from cuda_selection_criteria_b200 import synth

# id, n, seed, criterion, aux_bytes (-a), tau (-h), edge (inject duplicates + an empty sketch)
CASES = [
    dict(id="smh_n400_t090", n=400, seed=1001, criterion="smh_a", aux_bytes=1024, tau=0.9, edge=False),
    dict(id="smh_n400_t070", n=400, seed=1001, criterion="smh_a", aux_bytes=1024, tau=0.7, edge=False),
    dict(id="smh_n400_t080", n=400, seed=1001, criterion="smh_a", aux_bytes=1024, tau=0.8, edge=False),
    dict(id="smh_n400_t095", n=400, seed=1001, criterion="smh_a", aux_bytes=1024, tau=0.95, edge=False),
    dict(id="smh_n300_m32_t085", n=300, seed=7, criterion="smh_a", aux_bytes=256, tau=0.85, edge=True),
    dict(id="hlla_n400_p10_t090", n=400, seed=1003, criterion="hll_a", aux_bytes=1024, tau=0.9, edge=False),
    dict(id="hlla_n300_p8_t080", n=300, seed=1003, criterion="hll_a", aux_bytes=256, tau=0.8, edge=True),
    dict(id="hllan_n400_p10_t090", n=400, seed=1003, criterion="hll_an", aux_bytes=1024, tau=0.9, edge=False),
    dict(id="hllan_n300_p8_t080", n=300, seed=1003, criterion="hll_an", aux_bytes=256, tau=0.8, edge=True),
    # CB only (BASELINE config 2), emulated on the reference with -c smh_a -a 8 and identical
    # one-bucket .smh1 files, so smh_a always passes
    dict(id="cb_n300_t090", n=300, seed=1001, criterion="cb", aux_bytes=8, tau=0.9, edge=True),
]


def build_inputs(case) -> dict:
    n, p = case["n"], 14
    plan = synth.make_plan(n, case["seed"])
    regs = synth.hll(plan, p).copy()
    crit = case["criterion"]
    aux = None
    if crit == "smh_a":
        aux = synth.smh(plan, case["aux_bytes"] // 8).copy()
    elif crit in ("hll_a", "hll_an"):
        pa = case["aux_bytes"].bit_length() - 1
        aux = synth.hll(plan, pa, tag=synth.TAG_AUX_HLL).copy()
    if case["edge"]:
        # exact duplicates (cardinality ties -> std::sort tie order) and an empty sketch (e == 0 rows)
        regs[5] = regs[4]; regs[60] = regs[4]
        regs[7] = 0
        if aux is not None:
            aux[5] = aux[4]; aux[60] = aux[4]
            if crit != "smh_a":
                aux[7] = 0
    names = [f"g{i:05d}.fna.gz" for i in range(n)]
    return dict(regs=regs, aux=aux, names=names, p=p)


def digest(data) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(data["regs"]).tobytes())
    if data["aux"] is not None:
        h.update(np.ascontiguousarray(data["aux"]).tobytes())
    return h.hexdigest()
