"""synth-v1: deterministic synthetic sketches with controlled pairwise overlap.

Genomes come in clusters of geometric size; member = core(cluster) ∪ private(member) with
|core| log-uniform in [1e6, 8e6] 31-mers and target Jaccard(member, core) ~ U(0.6, 0.995).
The plan and the threshold tables are built here with numpy (Philox); the registers / buckets
are then drawn by integer-only code in the library (csrc/synth.cuh), on the host or on the
device with bit-identical results.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

TAG_PRIMARY, TAG_AUX_HLL, TAG_SMH = 1, 2, 3
_U64_MAX = np.uint64(0xFFFFFFFFFFFFFFFF)


@dataclass
class Plan:
    n: int
    seed: int
    cluster: np.ndarray    # int32 [n], non-decreasing
    n_core: np.ndarray     # float64 [n_clusters]
    n_priv: np.ndarray     # float64 [n]
    j_star: np.ndarray     # float64 [n] target Jaccard(member, core)

    @property
    def n_clusters(self) -> int:
        return int(self.n_core.size)

    def head(self, k: int) -> "Plan":
        """The first k genomes as their own plan (same draws: streams are keyed by genome / cluster id)."""
        ncl = int(self.cluster[k - 1]) + 1 if k else 0
        return Plan(k, self.seed, self.cluster[:k].copy(), self.n_core[:ncl].copy(), self.n_priv[:k].copy(),
                    self.j_star[:k].copy())


def make_plan(n: int, seed: int, mean_cluster: float = 20.0, core_range=(1e6, 8e6), j_range=(0.6, 0.995)) -> Plan:
    rng = np.random.Generator(np.random.Philox(seed))
    sizes = []
    tot = 0
    while tot < n:
        s = int(rng.geometric(1.0 / mean_cluster))
        s = min(s, n - tot)
        sizes.append(s)
        tot += s
    ncl = len(sizes)
    cluster = np.repeat(np.arange(ncl, dtype=np.int32), sizes)
    lo, hi = np.log(core_range[0]), np.log(core_range[1])
    n_core = np.exp(rng.uniform(lo, hi, ncl))
    j_star = rng.uniform(j_range[0], j_range[1], n)
    n_priv = n_core[cluster] * (1.0 / j_star - 1.0)
    return Plan(n, seed, cluster, n_core, n_priv, j_star)


def thresholds(card: np.ndarray, p: int) -> np.ndarray:
    """uint64 [len(card)][64]: T[k] = floor(2^64 · exp(-(card/2^p)·2^-k)), saturated; T[k>=64-p+1] = max."""
    lam = np.asarray(card, np.float64)[:, None] / float(1 << p)
    k = np.arange(64, dtype=np.float64)[None, :]
    x = np.exp(-lam * np.exp2(-k)) * 18446744073709551616.0
    x = np.minimum(x, 18446744073709549568.0)      # largest double below 2^64
    T = x.astype(np.uint64)
    T[:, 64 - p + 1:] = _U64_MAX
    T[np.asarray(card) <= 0] = _U64_MAX           # empty part: every register 0
    return np.ascontiguousarray(T)


def ranges(card: np.ndarray, m: int) -> np.ndarray:
    """uint64 [len(card)]: buckets uniform below 2^33/(card/m); 0 = empty part."""
    c = np.asarray(card, np.float64)
    per = np.maximum(1.0, c / m)
    r = np.floor(8589934592.0 / per).astype(np.uint64)
    r[c <= 0] = 0
    return np.ascontiguousarray(r)


def _alloc(shape, dtype, device):
    if device is None:
        a = np.empty(shape, dtype)
        return a, a.ctypes.data
    import torch
    tdt = {np.uint8: torch.uint8, np.uint64: torch.int64}[dtype]  # int64 storage viewed as uint64 bits
    t = torch.empty(shape, dtype=tdt, device=f"cuda:{device}")
    return t, t.data_ptr()


def hll(plan: Plan, p: int, tag: int = TAG_PRIMARY, device: int | None = None):
    """uint8 [n][2^p] registers; numpy on the host (device=None) or a CUDA torch tensor."""
    L = _lib.lib()
    tc = thresholds(plan.n_core, p)
    tp = thresholds(plan.n_priv, p)
    out, ptr = _alloc((plan.n, 1 << p), np.uint8, device)
    cl = np.ascontiguousarray(plan.cluster, np.int32)
    _lib.check(L.selb200_synth_hll(int(device is not None), int(device or 0), plan.n, p, cl.ctypes.data,
                                   plan.n_clusters, tc.ctypes.data, tp.ctypes.data, C.c_uint64(plan.seed),
                                   C.c_uint32(tag), ptr))
    return out


def smh(plan: Plan, m: int, tag: int = TAG_SMH, device: int | None = None):
    """uint64 [n][m] SuperMinHash buckets (torch: int64 storage holding the same bits)."""
    L = _lib.lib()
    rc = ranges(plan.n_core, m)
    rp = ranges(plan.n_priv, m)
    out, ptr = _alloc((plan.n, m), np.uint64, device)
    cl = np.ascontiguousarray(plan.cluster, np.int32)
    _lib.check(L.selb200_synth_smh(int(device is not None), int(device or 0), plan.n, m, cl.ctypes.data,
                                   plan.n_clusters, rc.ctypes.data, rp.ctypes.data, C.c_uint64(plan.seed),
                                   C.c_uint32(tag), ptr))
    return out
