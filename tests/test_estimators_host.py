"""CPU-only: the product's fp64 arithmetic (csrc/estimators.cuh: Ertl MLE, CB, hll_a, hll_an, Jaccard, sigma) compiled
for the host with contraction off and held bit-for-bit against the oracle (oracle/oracle.cpp, the restatement of
sketch/include/sketch/hll.h:628-688 and include/criteria_sketch.hpp:7-64 pinned to the reference's goldens), plus the
property the early exit of the MLE rests on: a stopped iteration never changes a decision."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_api as O
from cuda_selection_criteria_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def est(tmp_path_factory):
    so = tmp_path_factory.mktemp("est") / "libest.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                    os.path.join(ROOT, "tests", "emul", "estimators_host.cpp"), "-o", str(so)], check=True)
    L = C.CDLL(str(so))
    L.est_ertl_mle.restype = C.c_double
    L.est_ertl_mle.argtypes = [C.c_void_p, C.c_int]
    L.est_ertl_mle_stride.restype = C.c_double
    L.est_ertl_mle_stride.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.est_ertl_mle_stopj.restype = C.c_double
    L.est_ertl_mle_stopj.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint64, C.POINTER(C.c_int)]
    L.est_sigma_p.restype = C.c_float
    L.est_sigma_p.argtypes = [C.c_int]
    L.est_cb.argtypes = [C.c_double, C.c_uint64, C.c_uint64]
    L.est_hll_a.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_double, C.c_float]
    L.est_hll_an.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_double, C.c_float, C.c_int]
    L.est_hll_surely_fails.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
    L.est_hll_exact.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint64, C.c_float, C.c_int]
    L.est_jaccard.restype = C.c_double
    L.est_jaccard.argtypes = [C.c_uint64, C.c_uint64, C.c_double]
    return L


def hist64(regs):
    return np.bincount(regs, minlength=64).astype(np.uint32)


def bits(x):
    return np.float64(x).view(np.int64)


def test_ertl_mle_bit_exact_vs_oracle(est):
    ora = O.lib()
    checked = 0
    for p, n, seed in ((14, 300, 5), (10, 200, 6), (8, 200, 7), (4, 100, 8), (12, 100, 9)):
        plan = synth.make_plan(n, seed)
        regs = synth.hll(plan, p) if p == 14 else synth.hll(plan, p, synth.TAG_AUX_HLL)
        for g in range(n):
            h = hist64(regs[g])
            assert bits(est.est_ertl_mle(h.ctypes.data, p)) == bits(ora.oracle_ertl_mle(h.ctypes.data, p))
            checked += 1
        for g in range(0, n - 1, 3):       # unions of neighbours (cluster mates and strangers alike)
            h = hist64(np.maximum(regs[g], regs[g + 1]))
            t = est.est_ertl_mle(h.ctypes.data, p)
            assert bits(t) == bits(ora.oracle_ertl_mle(h.ctypes.data, p))
            assert bits(t) == bits(O.union_size(regs[g], regs[g + 1], p))
            checked += 1
    # corner histograms: empty sketch, one register set, everything saturated, everything at one value
    for p in (4, 9, 14):
        m, q = 1 << p, 64 - p
        for h in (np.array([m] + [0] * 63), np.array([m - 1, 1] + [0] * 62), np.eye(64, dtype=np.int64)[q + 1] * m,
                  np.eye(64, dtype=np.int64)[q] * m, np.eye(64, dtype=np.int64)[7] * m):
            h = h.astype(np.uint32)
            a, b = est.est_ertl_mle(h.ctypes.data, p), ora.oracle_ertl_mle(h.ctypes.data, p)
            assert bits(a) == bits(b) or (np.isinf(a) and np.isinf(b))
            checked += 1
    assert checked > 1200


def test_strided_histogram_is_the_same_estimate(est):
    """The hll filter keeps its histogram column-interleaved in shared memory (stride 64)."""
    plan = synth.make_plan(50, 11)
    regs = synth.hll(plan, 10, synth.TAG_AUX_HLL)
    for g in range(50):
        h = hist64(regs[g])
        wide = np.zeros(64 * 64, np.uint32)
        wide[::64] = h
        assert bits(est.est_ertl_mle_stride(wide.ctypes.data, 10, 64)) == bits(est.est_ertl_mle(h.ctypes.data, 10))


def test_early_exit_never_changes_a_decision(est):
    """DESIGN.md 'Early exit of the Ertl MLE': x only grows during the secant iteration, so J at the running bound
    is an upper bound of the final J.  Stopped => the full iteration fails J >= tau (and is outside the near-tau
    window); not stopped => the value is bit-for-bit the full iteration's."""
    plan = synth.make_plan(400, 21)
    regs = synth.hll(plan, 14)
    cards = np.array([O.cardinality(regs[g], 14) for g in range(400)])
    e = cards.astype(np.uint64)
    rng = np.random.default_rng(3)
    stopped_n = kept_n = 0
    for tau32 in (np.float32(0.9), np.float32(0.5), np.float32(0.99)):
        tau = float(tau32)
        for it in range(600):
            a, b = rng.integers(0, 400, 2).tolist()
            if it & 1:                                             # every other pair: two members of one cluster
                mates = np.flatnonzero(plan.cluster == plan.cluster[a])
                b = int(rng.choice(mates))
            if a == b:
                continue
            a, b = sorted((a, b), key=lambda g: cards[g])
            h = hist64(np.maximum(regs[a], regs[b]))
            full = est.est_ertl_mle(h.ctypes.data, 14)
            st = C.c_int(0)
            t = est.est_ertl_mle_stopj(h.ctypes.data, 14, tau, int(e[a]), int(e[b]), C.byref(st))
            j_full = est.est_jaccard(int(e[a]), int(e[b]), full)
            if st.value:
                stopped_n += 1
                assert t <= full                                   # a lower bound of the estimate
                assert j_full < tau - 1e-6 * abs(tau)              # neither emitted nor near
            else:
                kept_n += 1
                assert bits(t) == bits(full)
    assert stopped_n > 100 and kept_n > 100


def test_criteria_bit_exact_vs_oracle(est):
    ora = O.lib()
    rng = np.random.default_rng(17)
    for p in range(4, 15):
        assert np.float32(est.est_sigma_p(p)).view(np.int32) == np.float32(ora.oracle_sigma(p)).view(np.int32)
    n_true = n_false = 0
    for tau32 in (np.float32(0.9), np.float32(0.7), np.float32(0.95), np.float32(0.5)):
        tau = float(tau32)                                          # (double)(float) tau, as the reference compares
        for _ in range(1500):
            e2 = int(rng.integers(1, 10_000_000))
            e1 = int(e2 * rng.uniform(0.3, 1.0))
            # unions from "identical" to "disjoint", with the decision boundary well sampled
            t = (e1 + e2) / (1.0 + rng.uniform(0.2, 1.0)) * rng.uniform(0.97, 1.03)
            assert bool(est.est_cb(tau, e1, e2)) == bool(ora.oracle_cb(tau32, e1, e2))
            for p in (4, 5, 6, 7, 8, 10, 12):
                zs = np.float32(1.96) * np.float32(est.est_sigma_p(p))      # float product (criteria_sketch.hpp:29,32,40)
                a = bool(est.est_hll_a(tau, e1, e2, t, zs))
                assert a == bool(ora.oracle_hll_a(tau32, e1, e2, t, p, np.float32(1.96)))
                for order_n in (1, 2, 3):
                    b = bool(est.est_hll_an(tau, e1, e2, t, zs, order_n))
                    assert b == bool(ora.oracle_hll_an(tau32, e1, e2, t, p, np.float32(1.96), order_n))
                n_true += a
                n_false += not a
    assert n_true > 1000 and n_false > 1000


@pytest.mark.parametrize("p_aux,tau", [(10, 0.9), (8, 0.9), (6, 0.8), (12, 0.95), (9, 0.7)])
def test_hll_bound_never_rejects_a_passing_pair(est, p_aux, tau):
    """Pass A of the hll plane filter (selb::hll_surely_fails on upper bounds of the union's harmonic sum and empty count)
    against the exact decision (Ertl MLE + hll_a / hll_an): with all registers read, and after half / three quarters of them
    with the rest replaced by the smaller of the two genomes' own tail sums.  It may only discard pairs the exact decision
    discards; and it has to discard most strangers, or the two-pass filter gains nothing."""
    n = 500
    plan = synth.make_plan(n, 300 + p_aux)
    regs = synth.hll(plan, 14)
    aux = synth.hll(plan, p_aux, synth.TAG_AUX_HLL).astype(np.int64)
    cards = np.array([O.cardinality(regs[g], 14) for g in range(n)])
    order = np.argsort(cards, kind="stable")
    e = cards[order].astype(np.uint64)
    A = aux[order]
    m = 1 << p_aux
    tau32 = np.float32(tau)
    zs = np.float32(1.96) * np.float32(O.lib().oracle_sigma(p_aux))
    w = np.where(A > 0, np.exp2(-A.astype(np.float64)), 0.0)
    cuts = sorted({m, (3 * m) // 8, m // 2, (3 * m) // 4})      # all registers, and the three checkpoints of pass A
    tails = {c: (w[:, c:].sum(1), (A[:, c:] == 0).sum(1)) for c in cuts}
    tested = rejected_full = strangers = strangers_rejected_half = passing = 0
    cl = plan.cluster[order]
    for an in (0, 1):
        for i in range(0, n - 1, 3):
            for k in range(i + 1, n):
                if not est.est_cb(float(tau32), int(e[i]), int(e[k])):
                    break
                U = np.maximum(A[i], A[k])
                h = hist64(U.astype(np.uint8))
                exact = est.est_hll_exact(an, h.ctypes.data, p_aux, float(tau32), int(e[i]), int(e[k]), zs, 1)
                wu = np.where(U > 0, np.exp2(-U.astype(np.float64)), 0.0)
                passing += exact
                for c in cuts:
                    z = wu[:c].sum() + min(tails[c][0][i], tails[c][0][k])
                    c0 = (U[:c] == 0).sum() + min(tails[c][1][i], tails[c][1][k])
                    rej = est.est_hll_surely_fails(an, tau32, zs, 1, float(m), float(e[i]), float(e[k]),
                                                   float(np.float32(z) * np.float32(1.000001)), float(c0))
                    assert not (rej and exact), (an, i, k, c)
                    tested += 1
                    if c == m:
                        rejected_full += rej
                    if c == m // 2 and cl[i] != cl[k]:
                        strangers += 1
                        strangers_rejected_half += rej
    assert tested > 3000 and passing > 10
    assert rejected_full > 0.9 * (tested / len(cuts) - passing) - 5
    if p_aux >= 9 and tau >= 0.9:
        # strangers of equal size e: z_ub = 1.5 z_union after half the registers, so t_lb = 0.62 t = 1.23 e and k_mas = 0.73 —
        # decided at the first checkpoint for tau = 0.9, only at the second (k_mas = 0.47) for lower thresholds
        assert strangers_rejected_half > 0.99 * strangers
