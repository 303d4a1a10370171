"""CPU-only: the CB + smh_a chain of a run (k_cb_bounds, k_rowblock_span, k_tile_table, k_smh_signatures,
k_tile_filter_smh, k_smh_verify) compiled as host code from the .inl sources of the GPU build and run on the warp
emulator (tests/emul/), held against the oracle decision by decision: the CB band of every row
(src/selection.cpp:278-283, include/criteria_sketch.hpp:45-49), the number of pairs inside it, and the set of pairs
with an equal LSH band (criteria_sketch.hpp:66-81) — the bit-exact gate of the north star, without a GPU."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_api as O
from cuda_selection_criteria_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("emulf") / "emul_filter"
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-ffp-contract=off", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "emul", "emul_filter.cpp"), "-o", str(out)], check=True)
    return str(out)


def oracle_decisions(e, aux_sorted, tau32, n_rows, n_bands):
    """Row by row as the reference loops: skip e_k == 0, stop the row at the first CB failure, then smh_a."""
    ora = O.lib()
    n, m = aux_sorted.shape
    lo = np.zeros(n, np.int32)
    hi = np.zeros(n, np.int32)
    pairs = []
    p_cb = 0
    rows = [np.ascontiguousarray(aux_sorted[g]) for g in range(n)]
    zeros = int(np.count_nonzero(e == 0))
    for i in range(n):
        lo[i] = max(i + 1, zeros)
        k = lo[i]
        while k < n and ora.oracle_cb(tau32, int(e[i]), int(e[k])):
            p_cb += 1
            if ora.oracle_smh_a(rows[i].ctypes.data, rows[k].ctypes.data, m, n_rows, n_bands):
                pairs.append((i, k))
            k += 1
        hi[i] = k - 1
    return lo, hi, p_cb, pairs


def run_emulated(exe, tmp_path, e, aux_sorted, tau32, n_rows, n_bands, n_shards, grid, form="tiles"):
    n, m = aux_sorted.shape
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<5i", n, m, n_rows, n_bands, int(np.count_nonzero(e == 0))))
        f.write(struct.pack("<d", float(tau32)))          # tau = (double)(float) threshold, selection.cpp:81
        f.write(np.ascontiguousarray(e, np.uint64).tobytes())
        f.write(np.ascontiguousarray(aux_sorted, np.uint64).tobytes())
    r = subprocess.run([exe, str(inp), str(outp), str(n_shards), str(grid)] + form.split(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(outp, "rb").read()
    p_cb, tiles, cand, npairs = struct.unpack_from("<4q", raw, 0)
    off = 32
    lo = np.frombuffer(raw, np.int32, n, off); off += 4 * n
    hi = np.frombuffer(raw, np.int32, n, off); off += 4 * n
    pr = np.frombuffer(raw, np.uint32, 2 * npairs, off).reshape(-1, 2)
    return lo, hi, p_cb, tiles, cand, [tuple(x) for x in pr.tolist()]


@pytest.mark.parametrize("form", ["join", "join 9", "tiles"])     # "join 9": 9 signature bits in the bucket key (coarser buckets)
@pytest.mark.parametrize("n,seed,tau,m_aux,n_shards,grid", [
    (700, 41, 0.9, 128, 1, 3),        # the benchmark's shape: 16 bands x 8 rows, 8 signature words
    (600, 42, 0.75, 128, 2, 2),       # 32 bands x 4 rows: two chunks of signature words per tile; two shards
    (500, 43, 0.95, 64, 1, 5),        # 4 bands x 16 rows; more CTAs than some shards have tiles
    (300, 44, 0.9, 4, 3, 1),          # 2 bands x 2 rows: one signature word; three shards, one CTA
    (250, 45, 0.9, 1, 1, 2),          # 1 band x 1 row: the odd band count leaves a pad half that must never match
])
def test_cb_and_smh_a_decisions_on_the_emulator(exe, tmp_path, n, seed, tau, m_aux, n_shards, grid, form):
    tau32 = np.float32(tau)
    plan = synth.make_plan(n, seed)
    regs = synth.hll(plan, 14)
    smh = synth.smh(plan, m_aux)
    cards = np.array([O.cardinality(regs[g], 14) for g in range(n)])
    cards[::97] = 0.0                                      # a few empty sketches: e == 0 columns are skipped
    order = np.argsort(cards, kind="stable")
    e = cards[order].astype(np.uint64)                     # size_t e = card (selection.cpp:275,280)
    aux_sorted = np.ascontiguousarray(smh[order])
    n_bands, n_rows = O.band_params(m_aux, tau32)
    assert n_bands * n_rows == m_aux
    lo, hi, p_cb, tiles, cand, pairs = run_emulated(exe, tmp_path, e, aux_sorted, tau32, n_rows, n_bands, n_shards, grid, form)
    olo, ohi, op_cb, opairs = oracle_decisions(e, aux_sorted, tau32, n_rows, n_bands)
    rows_with_band = ohi >= olo
    assert np.array_equal(lo[rows_with_band], olo[rows_with_band]) and np.array_equal(hi[rows_with_band], ohi[rows_with_band])
    assert np.all(hi[~rows_with_band] < lo[~rows_with_band])
    assert p_cb == op_cb                                   # CB decisions bit-exact
    assert pairs == opairs                                 # smh_a decisions bit-exact, nothing lost by the 16-bit pre-filter
    assert len(opairs) > 20 and cand >= len(pairs) and tiles > 0


# ----------------------------------------------------------------------------------------------------------------
# hll_a / hll_an: k_aux_planes_quad + k_aux_range + the two-pass plane filter (k_tile_filter_hll_bound<AN> ->
# k_hll_verify<AN>), the single-pass plane filter k_tile_filter_hll_planes<AN> and k_tile_filter_hll<AN> (byte form; the
# histogram helpers are the emulator's semantic versions) against the oracle's decision per pair
# ----------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def exe_hll(tmp_path_factory):
    out = tmp_path_factory.mktemp("emulh") / "emul_filter_hll"
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-ffp-contract=off", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "emul", "emul_filter_hll.cpp"), "-o", str(out)], check=True)
    return str(out)


def oracle_hll_decisions(e, aux_sorted, p_aux, tau32, an, order_n):
    ora = O.lib()
    n = aux_sorted.shape[0]
    rows = [np.ascontiguousarray(aux_sorted[g]) for g in range(n)]
    zeros = int(np.count_nonzero(e == 0))
    pairs, p_cb = [], 0
    z = np.float32(1.96)
    for i in range(n):
        k = max(i + 1, zeros)
        while k < n and ora.oracle_cb(tau32, int(e[i]), int(e[k])):
            p_cb += 1
            t = ora.oracle_union_size(rows[i].ctypes.data, rows[k].ctypes.data, p_aux)
            ok = (ora.oracle_hll_an(tau32, int(e[i]), int(e[k]), t, p_aux, z, order_n) if an
                  else ora.oracle_hll_a(tau32, int(e[i]), int(e[k]), t, p_aux, z))
            if ok:
                pairs.append((i, k))
            k += 1
    return p_cb, pairs


@pytest.mark.parametrize("n,seed,tau,p_aux,criterion,form,n_shards,grid,odd", [
    # the default: pass A's fp32 bound (early exits after half / three quarters of the sketch) + pass B's exact decision
    (400, 51, 0.9, 8, "hll_a", "twopass", 1, 3, False),
    (400, 52, 0.85, 8, "hll_an", "twopass", 2, 2, False),
    (300, 53, 0.9, 10, "hll_a", "twopass", 1, 4, False),
    (300, 60, 0.9, 10, "hll_an", "twopass", 1, 4, False),
    (300, 54, 0.9, 6, "hll_an", "twopass", 1, 2, True),     # outliers: steps whose pairs do not share a 32-value window
    (300, 61, 0.7, 7, "hll_a", "twopass", 1, 2, True),
    (300, 62, 0.95, 12, "hll_a", "twopass", 2, 3, False),
    (300, 55, 0.9, 5, "hll_a", "bytes", 1, 2, False),       # p_aux < 6 has no bit planes: the byte form
    (300, 56, 0.8, 8, "hll_an", "bytes", 2, 3, True),
    # exact MLE for every pair of the band on the same planes (SELB200_HLLFILTER=onepass)
    (400, 57, 0.9, 8, "hll_a", "onepass", 1, 3, False),
    (300, 58, 0.85, 10, "hll_an", "onepass", 2, 2, False),
    (300, 59, 0.9, 6, "hll_a", "onepass", 1, 2, True),
])
def test_cb_and_hll_decisions_on_the_emulator(exe_hll, tmp_path, n, seed, tau, p_aux, criterion, form, n_shards, grid, odd):
    tau32 = np.float32(tau)
    an = criterion == "hll_an"
    plan = synth.make_plan(n, seed)
    regs = synth.hll(plan, 14)
    aux = synth.hll(plan, p_aux, synth.TAG_AUX_HLL).copy()
    if odd:
        aux[::7, 3] = 64 - p_aux + 1                       # the largest legal register, far above the rest
    cards = np.array([O.cardinality(regs[g], 14) for g in range(n)])
    cards[::89] = 0.0
    order = np.argsort(cards, kind="stable")
    e = cards[order].astype(np.uint64)
    aux_sorted = np.ascontiguousarray(aux[order])
    zs = np.float32(1.96) * np.float32(O.lib().oracle_sigma(p_aux))      # float product (criteria_sketch.hpp:29,32,40)
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<5i", n, p_aux, int(an), 1, int(np.count_nonzero(e == 0))))
        f.write(struct.pack("<d", float(tau32)))
        f.write(struct.pack("<f", float(zs)))
        f.write(e.tobytes())
        f.write(aux_sorted.tobytes())
    r = subprocess.run([exe_hll, str(inp), str(outp), form, str(n_shards), str(grid)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(outp, "rb").read()
    p_cb, npairs = struct.unpack_from("<2q", raw, 0)
    pairs = [tuple(x) for x in np.frombuffer(raw, np.uint32, 2 * npairs, 16).reshape(-1, 2).tolist()]
    op_cb, opairs = oracle_hll_decisions(e, aux_sorted, p_aux, tau32, an, 1)
    assert p_cb == op_cb
    assert pairs == opairs                                 # hll_a / hll_an decisions bit-exact (early-exit MLE included)
    assert 10 < len(opairs) < op_cb
    if form == "twopass":
        ncand = int(r.stdout.split("candidates=")[1].split()[0])
        assert len(opairs) <= ncand
        if not odd and p_aux >= 8:
            assert ncand < 2 * len(opairs) + 0.1 * op_cb   # the bound has to discard most failing pairs, or the second pass gains nothing


def test_hll_plane_histograms(exe_hll):
    """aux_plane_hist (the per-thread histogram step of k_hll_verify / k_tile_filter_hll_planes: subset counting on the
    quad plane layout) against hist[max(a[j], b[j])]++ on random sketch pairs: p_aux 6..12, every window, ranges of 1..24 values."""
    r = subprocess.run([exe_hll, "hist-check", "20261018", "4000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "hist-check: identical to the definition" in r.stdout
    assert int(r.stdout.split("bound sums valid on ")[1].split()[0]) > 1000     # pass A's twelve-value sums: valid upper bounds


_SMH_MUTATIONS = [
    ("wait leaves the group being read in flight",
     "if (item + 1 < staged) cp_async_wait<1>();", "if (item + 1 < staged) cp_async_wait<2>();"),
    ("refill queued before the barrier of the item",
     "        __syncthreads();\n        if (staged < n_items) { stage_next(); ++staged; }",
     "        if (staged < n_items) { stage_next(); ++staged; }\n        __syncthreads();"),
]


@pytest.mark.parametrize("name,old,new", _SMH_MUTATIONS, ids=[m[0] for m in _SMH_MUTATIONS])
def test_cp_async_model_catches_ring_mutations(tmp_path, name, old, new):
    """The emulator's cp.async model (destination poisoned when the copy is queued, written when wait_group retires
    its group) must reject an smh filter whose three-buffer ring is broken: the pair set then differs from the oracle's."""
    import shutil
    tree = tmp_path / "repo"
    shutil.copytree(os.path.join(ROOT, "tests", "emul"), tree / "tests" / "emul")
    shutil.copytree(os.path.join(ROOT, "cuda_selection_criteria_b200", "csrc"), tree / "cuda_selection_criteria_b200" / "csrc",
                    ignore=shutil.ignore_patterns("*.o", "*.so", "*.log", "variants"))
    src = tree / "cuda_selection_criteria_b200" / "csrc" / "kernels" / "filter_smh.inl"
    text = src.read_text()
    assert text.count(old) == 1, name
    src.write_text(text.replace(old, new))
    exe = tmp_path / "emul_filter_mut"
    subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-Wno-unknown-pragmas",
                    str(tree / "tests" / "emul" / "emul_filter.cpp"), "-o", str(exe)], check=True)
    n, m_aux, tau32 = 700, 128, np.float32(0.9)
    plan = synth.make_plan(n, 41)
    regs, smh = synth.hll(plan, 14), synth.smh(plan, m_aux)
    cards = np.array([O.cardinality(regs[g], 14) for g in range(n)])
    order = np.argsort(cards, kind="stable")
    e = cards[order].astype(np.uint64)
    aux_sorted = np.ascontiguousarray(smh[order])
    n_bands, n_rows = O.band_params(m_aux, tau32)
    _, _, op_cb, opairs = oracle_decisions(e, aux_sorted, tau32, n_rows, n_bands)
    for attempt in range(3):                   # the barrier mutation shows through thread timing: up to three runs
        _, _, p_cb, _, _, pairs = run_emulated(str(exe), tmp_path, e, aux_sorted, tau32, n_rows, n_bands, 1, 3)
        assert p_cb == op_cb                   # the band does not depend on the ring
        if pairs != opairs:
            return                             # the broken ring loses or invents pairs, and the check sees it
    assert pairs != opairs
