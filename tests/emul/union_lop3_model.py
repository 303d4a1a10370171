"""LOP3 / POPC per pair of the two counting forms of the plane union kernel on the pairs that reach the union pass of
the bench workload (synth-v1 seed 1002, tau = 0.9, CB + smh_a), from the value ranges of the genomes alone — the
counts follow from the warp-uniform group masks the kernel derives from `grange` (kernels/union_planes.inl):
    one-hot  (k_pair_hist_planes):                per 64 registers 4*NP + 16 + 34 * #8-groups in the range
    subsets  (k_pair_hist_planes<EpiSubsets<>>):  per 64 registers 4*NP + 2 * #8-groups + 16 * #4-groups
(NP = 5 planes when all values are below 32, else 6).  Uses the oracle for order and cardinalities, so it lives with
the test infrastructure:  python tests/emul/union_lop3_model.py [n_genomes]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api as O  # noqa: E402
from cuda_selection_criteria_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    p, m_aux, tau = 14, 128, float(np.float32(0.9))
    plan = synth.make_plan(n, 1002)
    regs, aux = synth.hll(plan, p), synth.smh(plan, m_aux)
    ora = O.select(regs, p, "smh_a", np.float32(0.9), aux=aux)
    order, e = ora["order"], np.floor(ora["cards_sorted"]).astype(np.uint64)
    bands, rows = int(ora["n_bands"]), int(ora["n_rows"])
    R, A = regs[order], aux[order].reshape(n, bands, rows)
    mn, mx = R.min(1).astype(int), R.max(1).astype(int)
    pairs = []
    for i in range(n):                                     # CB band of row i, then the LSH band test
        k = np.arange(i + 1, n)
        ok = (e[k] > 0) & (float(e[i]) / np.maximum(e[k], 1).astype(float) >= tau)
        if not ok.any():
            continue
        k = k[: np.nonzero(ok)[0].max() + 1]
        k = k[e[k] > 0]
        eq = (A[k] == A[i][None]).all(2).any(1)
        pairs += [(i, kk) for kk in k[eq]]
    pa = np.array(pairs)
    assert len(pa) == ora["stage"][2], (len(pa), ora["stage"])      # the oracle's count of aux-passing pairs
    lo, hi = np.maximum(mn[pa[:, 0]], mn[pa[:, 1]]), np.maximum(mx[pa[:, 0]], mx[pa[:, 1]])
    g8, g4 = (hi >> 3) - (lo >> 3) + 1, (hi >> 2) - (lo >> 2) + 1
    npl = np.where(np.minimum(lo >> 3, 4) == 0, 5, 6)
    steps = (1 << p) // 64 / 32                            # 64-register steps per lane and pair
    onehot, subsets = steps * (4 * npl + 16 + 34 * g8), steps * (4 * npl + 2 * g8 + 16 * g4)
    # the kernel as shipped: one mask of four counted directly (PL_DIRECT=1: 14 LOP3 and 5 POPC per group of four and step)
    # and, per 2048-register step, only the groups up to the larger of the two genomes' maxima over that eighth (gtop)
    blk = R.reshape(n, 8, -1).max(2).astype(int)                                   # per-eighth maxima
    top = np.maximum(blk[pa[:, 0]], blk[pa[:, 1]])                                 # [pairs][8]
    g4b = (top >> 2) - (lo >> 2)[:, None] + 1
    g8b = (top >> 3) - (lo >> 3)[:, None] + 1
    shipped = (4 * npl[:, None] + 2 * g8b + 14 * g4b).sum(1) * steps / 8
    shipped_popc = (5 * g4b).sum(1) * steps / 8
    print(f"n={n}: {len(pa)} pairs reach the union pass; value range lo {lo.mean():.2f} .. hi {hi.mean():.2f}, "
          f"8-groups {g8.mean():.2f}, 4-groups {g4.mean():.2f}")
    print(f"LOP3 per pair and lane: one-hot {onehot.mean():.0f}, subsets {subsets.mean():.0f} "
          f"(x{subsets.mean() / onehot.mean():.3f}); all four 8-groups: {steps * 172:.0f}")
    print(f"shipped subset kernel (direct mask + per-eighth tops): LOP3 {shipped.mean():.0f}, POPC {shipped_popc.mean():.0f}, "
          f"4-groups per step {g4b.mean():.2f}")
    print(f"POPC per pair and lane (loop): one-hot {(steps * 8 * g8).mean():.0f}, subsets {(steps * 4 * g4).mean():.0f}")


if __name__ == "__main__":
    main()
