"""CPU-only: a whole smh_a selection — load kernels, CB band, tile list, signatures, pre-filter, exact verification,
bit-plane union (+ byte kernel for wide pairs), Ertl-MLE estimate, emit, print-order sort — compiled as host code from
the .inl sources of the GPU build and run launch by launch on the warp emulator (tests/emul/emul_run.cpp), held against
the oracle's output for the same sketches: cardinalities and order, stage counts, the pair list in the reference's
print order and the Jaccard values.  What the -m gpu parity tests check on a B200, without the B200 (and without
the TMA / mbarrier protocol, which only the GPU can exercise)."""
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
    out = tmp_path_factory.mktemp("emulr") / "emul_run"
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-ffp-contract=off", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "emul", "emul_run.cpp"), "-o", str(out)], check=True)
    return str(out)


@pytest.mark.parametrize("n,seed,tau,m_aux,outliers,union_form,smh_form", [
    (130, 61, 0.9, 128, False, "planes", "join"),
    (140, 62, 0.9, 64, True, "planes", "join"),          # a few registers far above the rest: wide pairs take the byte kernel
    (120, 63, 0.9, 128, True, "subsets", "join"),        # the default forms: subset counting, equality join
    (150, 64, 0.7, 128, False, "subsets", "join"),       # 32 bands of 4 rows
    (130, 61, 0.9, 128, False, "subsets", "tiles"),      # the all-pairs tile filter + verify (SELB200_SMHFILTER=tiles)
    (130, 61, 0.9, 128, False, "subsets", "join screen"),  # estimate in two steps (screen, then the survivors): the cb route
])
def test_whole_smh_a_run_on_the_emulator(exe, tmp_path, n, seed, tau, m_aux, outliers, union_form, smh_form):
    tau32 = np.float32(tau)
    plan = synth.make_plan(n, seed)
    regs = synth.hll(plan, 14).copy()
    smh = synth.smh(plan, m_aux)
    if outliers:
        big = np.argsort(np.bincount(plan.cluster))[-4:]
        odd = np.concatenate([np.flatnonzero(plan.cluster == c)[:2] for c in big])
        regs[odd, 11] = 47
    regs[5] = 0                                            # an empty sketch: cardinality 0, skipped as e == 0
    n_bands, n_rows = O.band_params(m_aux, tau32)
    ora = O.select(regs, 14, "smh_a", tau32, aux=smh, threads=8)
    assert len(ora["i"]) > 10

    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<5i", n, 14, m_aux, n_rows, n_bands))
        f.write(struct.pack("<d", float(tau32)))
        f.write(np.ascontiguousarray(regs, np.uint8).tobytes())
        f.write(np.ascontiguousarray(smh, np.uint64).tobytes())
    r = subprocess.run([exe, str(inp), str(outp), union_form] + smh_form.split(), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(outp, "rb").read()
    p_cb, cand, p_aux, p_out, near, wide, tie, _ = struct.unpack_from("<8q", raw, 0)
    off = 64
    cards = np.frombuffer(raw, np.float64, n, off); off += 8 * n
    order = np.frombuffer(raw, np.int32, n, off); off += 4 * n
    keys = np.frombuffer(raw, np.uint64, p_out, off); off += 8 * p_out
    jac = np.frombuffer(raw, np.float64, p_out, off)

    # load: per-genome Ertl MLE bit-for-bit, same order
    assert tie == 0
    assert np.array_equal(order, ora["order"])
    assert np.array_equal(cards[order].view(np.int64), ora["cards_sorted"].view(np.int64))
    # run: stage counts, pair list in print order, Jaccard bits
    assert [p_cb, p_aux, p_out] == [int(x) for x in ora["stage"][1:4]]
    assert cand >= p_aux
    assert np.array_equal((keys >> np.uint64(32)).astype(np.int64), ora["i"].astype(np.int64))
    assert np.array_equal((keys & np.uint64(0xFFFFFFFF)).astype(np.int64), ora["k"].astype(np.int64))
    assert np.array_equal(jac.view(np.int64), ora["jaccard"].view(np.int64))
    assert (wide > 0) == outliers
