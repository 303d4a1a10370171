"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE (the checker, never the product)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None
CRIT = {"cb": 0, "smh_a": 1, "hll_a": 2, "hll_an": 3}


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True)
        L = C.CDLL(path)
        L.oracle_cardinality.restype = C.c_double
        L.oracle_cardinality.argtypes = [C.c_void_p, C.c_int]
        L.oracle_union_size.restype = C.c_double
        L.oracle_union_size.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_ertl_mle.restype = C.c_double
        L.oracle_ertl_mle.argtypes = [C.c_void_p, C.c_int]
        L.oracle_hist64.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.oracle_band_params.argtypes = [C.c_int, C.c_float, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_sort_order.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_cb.argtypes = [C.c_float, C.c_uint64, C.c_uint64]
        L.oracle_smh_a.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_hll_a.argtypes = [C.c_float, C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_float]
        L.oracle_hll_an.argtypes = [C.c_float, C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_float, C.c_int]
        L.oracle_sigma.restype = C.c_float
        L.oracle_sigma.argtypes = [C.c_int]
        L.oracle_select.restype = C.c_int64
        L.oracle_select.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                    C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
        L.oracle_near_count.restype = C.c_int64
        L.oracle_near_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_smh_size.restype = C.c_int
        L.oracle_smh_size.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def band_params(m, tau, cpu_variant=True):
    nb, nr = C.c_int(), C.c_int()
    lib().oracle_band_params(m, C.c_float(tau), int(cpu_variant), C.byref(nb), C.byref(nr))
    return nb.value, nr.value


def cardinality(regs: np.ndarray, p: int) -> float:
    r = np.ascontiguousarray(regs, np.uint8)
    return lib().oracle_cardinality(r.ctypes.data, p)


def union_size(a: np.ndarray, b: np.ndarray, p: int) -> float:
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().oracle_union_size(a.ctypes.data, b.ctypes.data, p)


def select(regs, p, criterion, tau, aux=None, aux_len=0, stored=None, z=1.96, order_n=1, n_rows=0, n_bands=0,
           threads=0, no_cb=False):
    """-> dict(i, k, jaccard, order, cards_sorted, stage=[P, P_cb, P_aux, P_out], near_i, near_k, near_jaccard);
    near_* = every evaluated pair with |J - tau| <= 1e-6*|tau| (emitted or not), in (i,k) order."""
    regs = np.ascontiguousarray(regs, np.uint8)
    n = regs.shape[0]
    crit = CRIT[criterion]
    a = None
    if crit == 1:
        a = np.ascontiguousarray(aux, np.uint64)
        aux_len = a.shape[1]
        if not (n_rows and n_bands):
            n_bands, n_rows = band_params(aux_len, tau, True)
    elif crit in (2, 3):
        a = np.ascontiguousarray(aux, np.uint8)
        aux_len = int(a.shape[1]).bit_length() - 1
    st = np.ascontiguousarray(stored, np.float64) if stored is not None else None
    cards = np.empty(n, np.float64); order = np.empty(n, np.int32)
    stage = np.zeros(4, np.int64)
    lib().oracle_set_no_cb(int(no_cb))
    cap = 1 << 16
    while True:
        lib().oracle_want_near(1)
        oi = np.empty(cap, np.int32); ok = np.empty(cap, np.int32); oj = np.empty(cap, np.float64)
        cnt = lib().oracle_select(n, p, regs.ctypes.data, st.ctypes.data if st is not None else None, crit, aux_len,
                                  a.ctypes.data if a is not None else None, C.c_float(tau), C.c_float(z), order_n,
                                  n_rows, n_bands, threads, cards.ctypes.data, order.ctypes.data, cap,
                                  oi.ctypes.data, ok.ctypes.data, oj.ctypes.data, None, stage.ctypes.data)
        if cnt <= cap:
            break
        cap = int(cnt)
    lib().oracle_set_no_cb(0)
    nn = lib().oracle_near_count()
    ni = np.empty(nn, np.int32); nk = np.empty(nn, np.int32); nj = np.empty(nn, np.float64)
    if nn:
        lib().oracle_near_copy(ni.ctypes.data, nk.ctypes.data, nj.ctypes.data)
    lib().oracle_want_near(0)
    return dict(i=oi[:cnt].copy(), k=ok[:cnt].copy(), jaccard=oj[:cnt].copy(), order=order, cards_sorted=cards,
                stage=stage.tolist(), n_rows=n_rows, n_bands=n_bands, near_i=ni, near_k=nk, near_jaccard=nj)


def format_lines(names, res):
    fi, fk = res["order"][res["i"]], res["order"][res["k"]]
    return [f"{names[a]} {names[b]} {j:f}" for a, b, j in zip(fi.tolist(), fk.tolist(), res["jaccard"].tolist())]


def ref_binary():
    path = os.path.join(ROOT, "oracle", "_ref", "selection")
    return path if os.path.exists(path) else None
