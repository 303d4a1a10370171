"""-m gpu: the CUDA path, called through the C-ABI, against the oracle and the reference goldens."""
import json
import os

import numpy as np
import pytest

import golden_cases as G
import oracle_api as O
import cuda_selection_criteria_b200 as S
from cuda_selection_criteria_b200 import sketch_io, synth
from cuda_selection_criteria_b200.selection import AUX_HLL, AUX_NONE, AUX_SMH

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
GOLD = os.path.join(HERE, "golden", "influenza")
REF_OUT = os.path.join(HERE, "golden", "ref_outputs")
REL_TOL = 1e-6   # north_star: Jaccard within 1e-6 relative; pair set and decisions bit-exact


def aux_kind_of(criterion):
    return {"cb": AUX_NONE, "smh_a": AUX_SMH, "hll_a": AUX_HLL, "hll_an": AUX_HLL}[criterion]


def run_gpu(regs, aux, criterion, tau, device=0, **kw):
    with S.Selection(device) as sel:
        sel.load(regs, aux, aux_kind_of(criterion))
        return sel.run(tau=np.float32(tau), criterion=criterion, **kw)


def compare(res, ora, tau):
    """Pair set bit-exact, same order, Jaccard within REL_TOL; stage counts identical.  Pairs within 1e-6 of tau
    (north_star: "listed separately") must be the ORACLE's near-tau pairs; only their tau decision is exempt from
    the exact check — every other pair, every Jaccard value and every stage count is still held to it."""
    near = list(zip(res.near_i.tolist(), res.near_k.tolist()))
    assert sorted(near) == list(zip(ora["near_i"].tolist(), ora["near_k"].tolist()))
    assert len(near) == res.stats["pairs_near"]
    near = set(near)
    got = dict(zip(zip(res.i.tolist(), res.k.tolist()), res.jaccard.tolist()))
    want = dict(zip(zip(ora["i"].tolist(), ora["k"].tolist()), ora["jaccard"].tolist()))
    assert [x for x in got if x not in near] == [x for x in want if x not in near]        # dicts keep list order
    assert np.array_equal(res.order, ora["order"])
    assert np.array_equal(res.cards_sorted.astype(np.uint64), ora["cards_sorted"].astype(np.uint64))
    for pr, j in got.items():
        if pr in want:
            assert abs(j - want[pr]) <= REL_TOL * abs(want[pr])
    nj = dict(zip(zip(ora["near_i"].tolist(), ora["near_k"].tolist()), ora["near_jaccard"].tolist()))
    for pr, j in zip(zip(res.near_i.tolist(), res.near_k.tolist()), res.near_jaccard.tolist()):
        assert abs(j - nj[pr]) <= REL_TOL * abs(nj[pr])
    st = res.stats
    assert st["pairs_total"] == ora["stage"][0]
    assert st["pairs_cb"] == ora["stage"][1]           # CB decisions bit-exact
    assert st["pairs_aux"] == ora["stage"][2]          # aux-criterion decisions bit-exact
    flipped = sum((pr in got) != (pr in want) for pr in near)
    assert abs(st["pairs_out"] - ora["stage"][3]) <= flipped
    if not near:
        assert st["pairs_out"] == ora["stage"][3]


@pytest.mark.parametrize("criterion,aux_bytes", [("smh_a", 512), ("smh_a", 32), ("hll_a", 256), ("hll_an", 256)])
def test_influenza_golden(gpu, criterion, aux_bytes):
    lines = S.run_filelist(os.path.join(GOLD, "test_influeza_filelist.txt"), tau=0.9, aux_bytes=aux_bytes,
                           criterion=criterion, device=gpu, base=GOLD)
    assert lines == open(os.path.join(GOLD, "results.txt")).read().splitlines()


@pytest.mark.parametrize("case", json.load(open(os.path.join(REF_OUT, "cases.json"))), ids=lambda c: c["id"])
def test_reference_binary_goldens(gpu, case):
    data = G.build_inputs(case)
    assert G.digest(data) == case["input_sha256"]
    res = run_gpu(data["regs"], data["aux"], case["criterion"], case["tau"], gpu)
    got = S.format_lines(data["names"], res)
    want = open(os.path.join(REF_OUT, case["id"] + ".txt")).read().splitlines()
    assert got == want
    ora = O.select(data["regs"], data["p"], case["criterion"], np.float32(case["tau"]), aux=data["aux"])
    compare(res, ora, case["tau"])


def test_cardinalities_and_union_estimates(gpu):
    plan = synth.make_plan(600, 11)
    regs = synth.hll(plan, 14)
    with S.Selection(gpu) as sel:
        sel.load(regs)
        cards, order = sel.order()
        want = np.array([O.cardinality(regs[i], 14) for i in range(regs.shape[0])])
        assert np.array_equal(order, np.argsort(want, kind="stable")) or np.array_equal(cards, np.sort(want))
        rel = np.abs(cards - want[order]) / want[order]
        assert rel.max() <= 1e-12
        rng = np.random.default_rng(5)
        a = rng.integers(0, 600, 4000).astype(np.int32)
        b = rng.integers(0, 600, 4000).astype(np.int32)
        t = sel.debug_union(a, b)
    want_t = np.array([O.union_size(regs[x], regs[y], 14) for x, y in zip(a, b)])
    assert (np.abs(t - want_t) / want_t).max() <= 1e-12


def test_synth_host_equals_device(gpu):
    plan = synth.make_plan(257, 99)
    for p, tag in [(14, synth.TAG_PRIMARY), (10, synth.TAG_AUX_HLL), (8, synth.TAG_AUX_HLL)]:
        h = synth.hll(plan, p, tag)
        d = synth.hll(plan, p, tag, device=gpu).cpu().numpy()
        assert np.array_equal(h, d)
    h = synth.smh(plan, 128)
    d = synth.smh(plan, 128, device=gpu).cpu().numpy().view(np.uint64)
    assert np.array_equal(h, d)


@pytest.mark.parametrize("criterion,tau,n", [("smh_a", 0.9, 3000), ("smh_a", 0.75, 1500), ("cb", 0.9, 1200),
                                             ("hll_a", 0.9, 2000), ("hll_an", 0.9, 2000), ("hll_a", 0.8, 900),
                                             ("hll_an", 0.7, 900)])
def test_synthetic_vs_oracle(gpu, criterion, tau, n):
    plan = synth.make_plan(n, 1002)
    regs = synth.hll(plan, 14)
    aux = None
    if criterion == "smh_a":
        aux = synth.smh(plan, 128)
    elif criterion in ("hll_a", "hll_an"):
        aux = synth.hll(plan, 10, synth.TAG_AUX_HLL)
    res = run_gpu(regs, aux, criterion, tau, gpu)
    ora = O.select(regs, 14, criterion, np.float32(tau), aux=aux, threads=8)
    assert len(ora["i"]) > 0
    compare(res, ora, tau)


@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_shards_union_equals_whole(gpu, n_shards):
    plan = synth.make_plan(2500, 21)
    regs = synth.hll(plan, 14)
    aux = synth.smh(plan, 128)
    with S.Selection(gpu) as sel:
        sel.load(regs, aux, AUX_SMH)
        whole = sel.run(tau=np.float32(0.85), criterion="smh_a")
        parts = [sel.run(tau=np.float32(0.85), criterion="smh_a", shard=s, n_shards=n_shards) for s in range(n_shards)]
    keys = np.concatenate([(p.i.astype(np.int64) << 32) | p.k for p in parts])
    jac = np.concatenate([p.jaccard for p in parts])
    o = np.argsort(keys, kind="stable")
    assert np.array_equal(keys[o], (whole.i.astype(np.int64) << 32) | whole.k)
    assert np.array_equal(jac[o], whole.jaccard)
    assert sum(p.stats["tiles_shard"] for p in parts) == whole.stats["tiles_total"]
    assert sum(p.stats["pairs_aux"] for p in parts) == whole.stats["pairs_aux"]


def test_device_resident_load_matches_host_load(gpu):
    import torch
    plan = synth.make_plan(1000, 3)
    regs_d = synth.hll(plan, 14, device=gpu)
    aux_d = synth.smh(plan, 128, device=gpu)
    with S.Selection(gpu) as sel:
        sel.load(regs_d, aux_d, AUX_SMH)
        a = sel.run(tau=np.float32(0.9), criterion="smh_a")
    b = run_gpu(regs_d.cpu().numpy(), aux_d.cpu().numpy().view(np.uint64), "smh_a", 0.9, gpu)
    assert np.array_equal(a.i, b.i) and np.array_equal(a.k, b.k) and np.array_equal(a.jaccard, b.jaccard)
    assert torch.cuda.is_available()


def test_edge_cases(gpu):
    plan = synth.make_plan(64, 5)
    regs = synth.hll(plan, 14)
    aux = synth.smh(plan, 128)
    # empty and tiny inputs
    for n in (0, 1, 2):
        r = run_gpu(regs[:n], aux[:n], "smh_a", 0.9, gpu)
        o = O.select(regs[:n], 14, "smh_a", np.float32(0.9), aux=aux[:n]) if n else None
        assert r.stats["pairs_total"] == n * (n - 1) // 2
        if o is not None:
            compare(r, o, 0.9)
    # tau > 1 selects nothing; tau = 0 passes CB for every pair
    assert run_gpu(regs, aux, "smh_a", 1.5, gpu).i.size == 0
    r0 = run_gpu(regs, None, "cb", 0.0, gpu)
    o0 = O.select(regs, 14, "cb", np.float32(0.0))
    compare(r0, o0, 0.0)
    assert r0.stats["pairs_cb"] == 64 * 63 // 2
    # all sketches identical: every cardinality ties, every pair has J = 1
    same = np.repeat(regs[:1], 40, axis=0)
    rs = run_gpu(same, np.repeat(aux[:1], 40, axis=0), "smh_a", 0.9, gpu)
    os_ = O.select(same, 14, "smh_a", np.float32(0.9), aux=np.repeat(aux[:1], 40, axis=0))
    compare(rs, os_, 0.9)
    assert rs.i.size == 40 * 39 // 2
    # smh band shape that does not tile the sketch: the reference's smh_a returns 0 for every pair
    with S.Selection(gpu) as sel:
        sel.load(regs, aux, AUX_SMH)
        assert sel.run(tau=np.float32(0.9), criterion="smh_a", n_rows=3, n_bands=5).i.size == 0
    # empty sketches (e == 0) among real ones
    z = regs.copy(); z[3] = 0; z[9] = 0
    rz = run_gpu(z, aux, "smh_a", 0.7, gpu)
    oz = O.select(z, 14, "smh_a", np.float32(0.7), aux=aux)
    compare(rz, oz, 0.7)


def test_errors_are_loud(gpu):
    plan = synth.make_plan(8, 5)
    regs = synth.hll(plan, 14).copy()
    with S.Selection(gpu) as sel:
        with pytest.raises(S.SelB200Error):
            sel.run()                                   # run before load
        bad = regs.copy(); bad[2, 100] = 52             # > 64-14+1
        with pytest.raises(S.SelB200Error):
            sel.load(bad)
        sel.load(regs)
        with pytest.raises(S.SelB200Error):
            sel.run(criterion="smh_a")                  # no SuperMinHash sketches loaded
        with pytest.raises(S.SelB200Error):
            sel.run(criterion="hll_a")


def test_other_primary_precisions(gpu):
    plan = synth.make_plan(300, 17)
    for p in (10, 12, 15, 16):      # 15, 16: an eighth of the sketch (the union kernel's per-step group limit) spans 2 / 4 steps
        regs = synth.hll(plan, p)
        r = run_gpu(regs, None, "cb", 0.9, gpu)
        o = O.select(regs, p, "cb", np.float32(0.9))
        compare(r, o, 0.9)


def test_streaming_load_equals_bulk_load(gpu):
    """selb200_load_begin/acquire/commit/end over pinned slots == one selb200_load_host call."""
    n = 9000          # > 2 chunks of 4096 rows
    plan = synth.make_plan(n, 44)
    regs = synth.hll(plan, 14)
    aux = synth.smh(plan, 128)
    stored = np.full(n, -1.0)
    stored[17] = 1234567.25                      # a trusted stored cardinality (hll.h:1138-1141)

    def fill(g0, cnt, r, st, ax):
        r[:] = regs[g0:g0 + cnt]
        st[:] = stored[g0:g0 + cnt]
        ax[:] = aux[g0:g0 + cnt]

    with S.Selection(gpu) as sel:
        sel.load_stream(n, 14, fill, AUX_SMH, 128)
        a = sel.run(tau=np.float32(0.9), criterion="smh_a")
    with S.Selection(gpu) as sel:
        sel.load(regs, aux, AUX_SMH, stored=stored)
        b = sel.run(tau=np.float32(0.9), criterion="smh_a")
    assert np.array_equal(a.order, b.order) and np.array_equal(a.cards_sorted, b.cards_sorted)
    assert 1234567.25 in a.cards_sorted
    assert np.array_equal(a.i, b.i) and np.array_equal(a.k, b.k) and np.array_equal(a.jaccard, b.jaccard)
    ora = O.select(regs, 14, "smh_a", np.float32(0.9), aux=aux, stored=stored, threads=8)
    compare(a, ora, 0.9)


def test_malformed_aux_hll_is_rejected(gpu):
    plan = synth.make_plan(40, 5)
    regs = synth.hll(plan, 14)
    aux = synth.hll(plan, 8, synth.TAG_AUX_HLL).copy()
    aux[3, 7] = 64 - 8 + 2
    with S.Selection(gpu) as sel:
        with pytest.raises(S.SelB200Error):
            sel.load(regs, aux, AUX_HLL)


def test_no_cb_mode_matches_time_smh_loop(gpu):
    """The "smh_a" loop of experiments/src/time_smh.cpp:229-257: no cardinality bound, e2 == 0 still skipped."""
    plan = synth.make_plan(700, 31)
    regs = synth.hll(plan, 14).copy()
    regs[11] = 0
    aux = synth.smh(plan, 128)
    with S.Selection(gpu) as sel:
        sel.load(regs, aux, AUX_SMH)
        r = sel.run(tau=np.float32(0.8), criterion="smh_a", no_cb=True)
    o = O.select(regs, 14, "smh_a", np.float32(0.8), aux=aux, no_cb=True)
    compare(r, o, 0.8)
    assert r.stats["pairs_cb"] == 700 * 699 // 2          # the empty genome sorts first, so it is never a second genome


def test_host_results_and_both_sort_paths(gpu):
    """params.host_results: the lists are already in pinned host memory when the run returns; sparse outputs
    take the row-bucket sort, dense ones (more than 4 pairs per genome) the radix sort — same order."""
    plan = synth.make_plan(1500, 77)
    regs, aux = synth.hll(plan, 14), synth.smh(plan, 128)
    with S.Selection(gpu) as sel:
        sel.load(regs, aux, AUX_SMH)
        for crit, tau in (("smh_a", 0.9), ("cb", 0.5), ("smh_a", 0.3)):
            res = sel.run(tau=np.float32(tau), criterion=crit)
            keys, jac = sel.result_host()
            assert np.array_equal(keys, (res.i.astype(np.uint64) << np.uint64(32)) | res.k.astype(np.uint64))
            assert np.array_equal(jac, res.jaccard)
            assert np.all(np.diff(keys.astype(np.int64)) > 0)          # strictly (i,k)-ordered
            ora = O.select(regs, 14, crit, np.float32(tau), aux=aux if crit != "cb" else None, threads=8)
            compare(res, ora, tau)
        dense = sel.run(tau=np.float32(0.5), criterion="cb")
        assert dense.i.size > 4 * 1500
        raw = sel.run(tau=np.float32(0.9), criterion="smh_a", sort_output=False)
        srt = sel.run(tau=np.float32(0.9), criterion="smh_a")
        k_raw = (raw.i.astype(np.int64) << 32) | raw.k
        o = np.argsort(k_raw, kind="stable")
        assert np.array_equal(k_raw[o], (srt.i.astype(np.int64) << 32) | srt.k) and np.array_equal(raw.jaccard[o], srt.jaccard)


def test_piecewise_device_load_equals_bulk_load(gpu):
    """selb200_load_device_begin / _rows / load_end: pieces declared in any order give the bulk result."""
    plan = synth.make_plan(1300, 9)
    regs_d = synth.hll(plan, 14, device=gpu)
    aux_d = synth.smh(plan, 128, device=gpu)
    import torch
    torch.cuda.synchronize()
    with S.Selection(gpu) as sel:
        sel.load(regs_d, aux_d, AUX_SMH)
        bulk = sel.run(tau=np.float32(0.85), criterion="smh_a")
        sel.load_device_begin(regs_d, aux_d, AUX_SMH)
        for g0, cnt in ((900, 400), (0, 333), (333, 567)):
            sel.load_device_rows(g0, cnt)
        sel.load_end()
        piece = sel.run(tau=np.float32(0.85), criterion="smh_a")
        assert np.array_equal(bulk.i, piece.i) and np.array_equal(bulk.k, piece.k)
        assert np.array_equal(bulk.jaccard, piece.jaccard) and np.array_equal(bulk.order, piece.order)
        # a row left undeclared is an error, not a silent zero
        sel.load_device_begin(regs_d, aux_d, AUX_SMH)
        sel.load_device_rows(0, 1000)
        with pytest.raises(S.SelB200Error):
            sel.load_end()


def test_wide_value_range_pairs_take_the_byte_kernel(gpu):
    """Genomes with a register far above the rest (value window wider than 32) make every pair they are in
    'wide': those pairs leave the bit-plane union kernel for the byte kernel, whole batches of them in a
    row.  Results stay the oracle's, run after run (the batch bookkeeping once lost its place here)."""
    plan = synth.make_plan(1500, 31)
    regs = synth.hll(plan, 14).copy()
    aux = synth.smh(plan, 128)
    sizes = np.bincount(plan.cluster)
    big = np.argsort(sizes)[-6:]                      # members of the six largest clusters
    odd = np.concatenate([np.flatnonzero(plan.cluster == c)[:3] for c in big])
    regs[odd, 7] = 45                                 # legal for p=14 (<= 51), far above everything else
    ora = O.select(regs, 14, "smh_a", np.float32(0.8), aux=aux, threads=8)
    assert len(ora["i"]) > 0
    with S.Selection(gpu) as sel:
        sel.load(regs, aux, AUX_SMH)
        for _ in range(12):
            compare(sel.run(tau=np.float32(0.8), criterion="smh_a"), ora, 0.8)
    res = run_gpu(regs, None, "cb", 0.97, gpu)        # dense list: long runs of wide pairs
    compare(res, O.select(regs, 14, "cb", np.float32(0.97), threads=8), 0.97)


@pytest.mark.parametrize("form", ["subsets", "planes"])
def test_union_mixed_value_ranges(gpu, monkeypatch, form):
    """Both counting steps of the plane union kernel (SELB200_UNION is read when a context is created; subsets is
    the default) on genomes whose registers start at different values, hold long runs of high values and reach the
    largest legal value: pairs whose values do not fit one 32-value window leave for the byte kernel, the rest is
    counted on the planes, and the result must not depend on which."""
    monkeypatch.setenv("SELB200_UNION", form)
    plan = synth.make_plan(1200, 77)
    regs = synth.hll(plan, 14).copy()
    aux = synth.smh(plan, 128)
    regs[::3] = np.maximum(regs[::3], 8)              # every third genome: smallest register 8 -> window base 8
    regs[5::50, ::8] = 30                             # 2048 high registers
    regs[7::50, ::40] = 29                            # 410 high registers
    regs[11::50, 100:140] = 51                        # the largest legal value, a run of neighbours
    for crit, tau, a in (("cb", 0.93, None), ("smh_a", 0.8, aux)):
        ora = O.select(regs, 14, crit, np.float32(tau), aux=a, threads=8)
        assert ora["stage"][2] > 1000
        with S.Selection(gpu) as sel:
            sel.load(regs, a, aux_kind_of(crit))
            for _ in range(3):
                compare(sel.run(tau=np.float32(tau), criterion=crit), ora, tau)


@pytest.mark.parametrize("criterion", ["hll_a", "hll_an"])
@pytest.mark.parametrize("p_aux", [4, 5, 6, 7, 9, 12])
def test_auxiliary_hll_precisions(gpu, criterion, p_aux):
    """Every auxiliary precision the reference's sigma table distinguishes (criteria_sketch.hpp:7-20: p = 4..7
    have their own constants) plus a large one; p_aux < 6 takes the byte filter, the rest the bit-plane
    filter with 2, 4, 16 and 128 words per plane."""
    plan = synth.make_plan(700, 100 + p_aux)
    regs = synth.hll(plan, 14)
    aux = synth.hll(plan, p_aux, synth.TAG_AUX_HLL)
    tau = 0.8
    res = run_gpu(regs, aux, criterion, tau, gpu)
    ora = O.select(regs, 14, criterion, np.float32(tau), aux=aux, threads=8)
    assert ora["stage"][2] > 0
    compare(res, ora, tau)


@pytest.mark.parametrize("criterion", ["cb", "smh_a"])
def test_pairs_within_1e6_of_tau_are_listed_separately(gpu, criterion):
    """tau chosen as the float nearest to the Jaccard of an actual pair puts that pair within 6e-8 (relative) of tau,
    below it or above it depending on the rounding: once on each side, plus a tau with several near pairs.  The near
    list must be the oracle's, an emitted near pair appears in both lists, and everything else stays exact."""
    plan = synth.make_plan(900, 4242)
    regs = synth.hll(plan, 14)
    aux = synth.smh(plan, 128) if criterion == "smh_a" else None
    base = O.select(regs, 14, criterion, np.float32(0.80), aux=aux, threads=8)
    js = base["jaccard"][(base["jaccard"] > 0.86) & (base["jaccard"] < 0.97)]
    assert js.size > 50
    above = [j for j in js if float(np.float32(j)) > j][:2]        # tau just above J: evaluated, near, NOT emitted
    below = [j for j in js if float(np.float32(j)) < j][:2]        # tau just below J: near AND emitted
    assert above and below
    sides = set()
    with S.Selection(gpu) as sel:
        sel.load(regs, aux, aux_kind_of(criterion))
        for j in above + below:
            tau = np.float32(j)
            ora = O.select(regs, 14, criterion, tau, aux=aux, threads=8)
            assert ora["near_i"].size >= 1
            res = sel.run(tau=tau, criterion=criterion)
            compare(res, ora, tau)
            emitted = set(zip(res.i.tolist(), res.k.tolist()))
            for pr, jn in zip(zip(res.near_i.tolist(), res.near_k.tolist()), res.near_jaccard.tolist()):
                assert (pr in emitted) == (jn >= float(tau))
                sides.add(jn >= float(tau))
    assert sides == {True, False}


def test_near_tau_list_grows_instead_of_truncating(gpu):
    """400 identical sketches at tau = 1: every pair has J = (2e - t)/t with e = trunc(t), i.e. 1 - 2 frac(t)/t — within
    1e-6 of tau and below it.  79 800 near pairs exceed the list's initial 65 536 entries: the run must grow the list
    and return all of them (it used to cut the list silently)."""
    plan = synth.make_plan(3, 5)
    one = synth.hll(plan, 14)[:1]
    regs = np.repeat(one, 400, axis=0)
    ora = O.select(regs, 14, "cb", np.float32(1.0), threads=8)
    assert ora["near_i"].size == 400 * 399 // 2 and ora["stage"][3] == 0
    res = run_gpu(regs, None, "cb", 1.0, gpu)
    assert res.stats["pairs_near"] == 79800 and res.near_i.size == 79800
    compare(res, ora, 1.0)


# ---- BASELINE.json configs at their own size (SURVEY.md §8d C2, C3, C5's criteria at n = 10k) ----------------
NCPU = os.cpu_count() or 8


def test_config_c2_cb_only_10k(gpu):
    """C2: n = 10 000, CB only, tau = 0.9 — 4.7 M pairs through the HLL-14 union, every one against the oracle."""
    plan = synth.make_plan(10_000, 1001)
    regs = synth.hll(plan, 14)
    res = run_gpu(regs, None, "cb", 0.9, gpu)
    ora = O.select(regs, 14, "cb", np.float32(0.9), threads=NCPU)
    assert ora["stage"][1] == ora["stage"][2] > 4_000_000
    compare(res, ora, 0.9)


def test_config_c3_tau_sweep_10k(gpu):
    """C3: n = 10 000, CB + smh_a (1 KiB, m = 128), all six thresholds of the sweep on one loaded context."""
    plan = synth.make_plan(10_000, 1001)
    regs = synth.hll(plan, 14)
    aux = synth.smh(plan, 128)
    with S.Selection(gpu) as sel:
        sel.load(regs, aux, AUX_SMH)
        for tau in (0.70, 0.75, 0.80, 0.85, 0.90, 0.95):
            res = sel.run(tau=np.float32(tau), criterion="smh_a")
            ora = O.select(regs, 14, "smh_a", np.float32(tau), aux=aux, threads=NCPU)
            assert (res.stats["n_bands"], res.stats["n_rows"]) == (ora["n_bands"], ora["n_rows"])
            assert len(ora["i"]) > 500
            compare(res, ora, tau)


@pytest.mark.parametrize("criterion", ["hll_a", "hll_an"])
@pytest.mark.parametrize("p_aux", [10, 8])
def test_config_c5_criteria_10k(gpu, criterion, p_aux):
    """C5's criteria and auxiliary sizes (-a 1024 and -a 256) on 10 000 genomes of its seed: every CB pair's hll_a /
    hll_an decision (P_aux) and the final list against the oracle."""
    plan = synth.make_plan(10_000, 1003)
    regs = synth.hll(plan, 14)
    aux = synth.hll(plan, p_aux, synth.TAG_AUX_HLL)
    res = run_gpu(regs, aux, criterion, 0.9, gpu)
    ora = O.select(regs, 14, criterion, np.float32(0.9), aux=aux, threads=NCPU)
    assert ora["stage"][1] > 4_000_000 and ora["stage"][3] > 500
    compare(res, ora, 0.9)
