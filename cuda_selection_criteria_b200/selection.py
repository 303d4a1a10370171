"""Host mirror of the reference's selection drivers over the C-ABI (include/selb200.h).

Mirrors src/selection.cpp:70-303 (flags -l -t -a -h -c) and src/selection_cuda.cpp:59-189
(-l -b -a -h): load the sketch files, hand the matrices to the CUDA library, print
``nameA nameB jaccard`` lines in the reference's order with std::to_string formatting.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib, sketch_io

CRITERIA = {"cb": 0, "smh_a": 1, "hll_a": 2, "hll_an": 3}
AUX_NONE, AUX_SMH, AUX_HLL = 0, 1, 2


def band_params(m: int, tau: float, cpu_variant: bool = True) -> tuple[int, int]:
    """(n_bands, n_rows) of the LSH band search: src/selection.cpp:258-267 (cpu_variant) or
    src/selection_cuda.cpp:119-128."""
    nb, nr = C.c_int(), C.c_int()
    _lib.check(_lib.lib().selb200_band_params(int(m), C.c_float(tau), int(cpu_variant), C.byref(nb), C.byref(nr)))
    return nb.value, nr.value


@dataclass
class SelectionResult:
    i: np.ndarray            # sorted positions, i < k
    k: np.ndarray
    jaccard: np.ndarray      # float64, the value the reference formats with std::to_string
    order: np.ndarray        # sorted position -> file-list index
    cards_sorted: np.ndarray
    stats: dict
    near_i: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    near_k: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    near_jaccard: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float64))

    def file_indices(self):
        return self.order[self.i], self.order[self.k]


def _is_cuda_tensor(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


class Selection:
    """One selection context on one GPU (one per process per device).

    `stream`: a cudaStream_t handle (e.g. ``torch.cuda.Stream().cuda_stream``) the library's kernels run
    on, so that torch copies / NCCL collectives issued on the same stream are ordered with them.  None or
    0 (the handle of the legacy default stream) = a non-blocking stream owned by the context: device
    tensors handed to load() must then be complete (synchronise first)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._L = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._L.selb200_create(int(device), C.c_void_p(stream or 0), C.byref(h)))
        self._h = h
        self.device = device
        self.n = 0
        self._keep = None  # borrowed device tensors must outlive the context's use of them

    def close(self):
        if getattr(self, "_h", None):
            self._L.selb200_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- load -----------------------------------------------------------------------------
    def load(self, regs, aux=None, aux_kind: int = AUX_NONE, aux_len: int = 0, stored=None, p: int | None = None):
        """regs: uint8 [n][2^p] (numpy = host copy, CUDA torch tensor = borrowed device memory).
        aux: uint64 [n][m] (AUX_SMH) or uint8 [n][2^p_aux] (AUX_HLL) of the same residency."""
        n = int(regs.shape[0])
        m = int(regs.shape[1]) if n or len(regs.shape) > 1 else 1 << 14
        if p is None:
            p = m.bit_length() - 1
        if (1 << p) != m:
            raise ValueError(f"register rows of {m} bytes are not a power of two")
        if aux_kind == AUX_SMH and aux is not None and not aux_len:
            aux_len = int(aux.shape[1])
        if aux_kind == AUX_HLL and aux is not None and not aux_len:
            aux_len = int(aux.shape[1]).bit_length() - 1
        st = None
        if stored is not None:
            st = np.ascontiguousarray(stored, dtype=np.float64)
        stp = st.ctypes.data if st is not None else None
        if _is_cuda_tensor(regs):
            if not regs.is_contiguous() or (aux is not None and not aux.is_contiguous()):
                raise ValueError("device tensors must be contiguous")
            self._keep = (regs, aux)
            _lib.check(self._L.selb200_load_device(self._h, n, p, regs.data_ptr(), stp, aux_kind, aux_len,
                                                   aux.data_ptr() if aux is not None else None))
        else:
            if hasattr(regs, "numpy"):  # pinned CPU torch tensor
                r_ptr, a_ptr = regs.data_ptr(), (aux.data_ptr() if aux is not None else None)
                self._keep = (regs, aux)
            else:
                r = np.ascontiguousarray(regs, dtype=np.uint8)
                a = None
                if aux is not None:
                    a = np.ascontiguousarray(aux, dtype=np.uint64 if aux_kind == AUX_SMH else np.uint8)
                self._keep = (r, a)
                r_ptr, a_ptr = r.ctypes.data, (a.ctypes.data if a is not None else None)
            _lib.check(self._L.selb200_load_host(self._h, n, p, r_ptr, stp, aux_kind, aux_len, a_ptr))
        self.n = n
        self.p = p
        self._order = None
        return self

    def load_device_begin(self, regs, aux=None, aux_kind: int = AUX_NONE, aux_len: int = 0):
        """Piecewise load of device matrices that are still being filled (CUDA torch tensors, borrowed):
        follow with load_device_rows(g0, count) for every piece — issued on the context's stream behind the
        copy / collective that completes those rows — and load_end()."""
        n, m = int(regs.shape[0]), int(regs.shape[1])
        p = m.bit_length() - 1
        if (1 << p) != m or not regs.is_contiguous() or (aux is not None and not aux.is_contiguous()):
            raise ValueError("register rows must be a power of two wide and the tensors contiguous")
        if aux_kind == AUX_SMH and aux is not None and not aux_len:
            aux_len = int(aux.shape[1])
        if aux_kind == AUX_HLL and aux is not None and not aux_len:
            aux_len = int(aux.shape[1]).bit_length() - 1
        self._keep = (regs, aux)
        _lib.check(self._L.selb200_load_device_begin(self._h, n, p, regs.data_ptr(), aux_kind, aux_len,
                                                     aux.data_ptr() if aux is not None else None))
        self.n, self.p, self._order = n, p, None
        return self

    def load_device_rows(self, g0: int, count: int):
        _lib.check(self._L.selb200_load_device_rows(self._h, int(g0), int(count)))

    def load_info(self) -> dict:
        """Register bytes the last host load moved over PCIe, and how many rows went packed / raw."""
        b, rp, rr = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self._L.selb200_load_info(self._h, C.byref(b), C.byref(rp), C.byref(rr)))
        return {"h2d_register_bytes": b.value, "rows_packed": rp.value, "rows_raw": rr.value}

    def load_device_rows_packed(self, g0: int, count: int, pieces, piece_rows: int | None = None):
        """Rows [g0, g0+count) arrived as packed pieces (selb200_nib4_pack_piece) in device memory: `pieces` is a CUDA
        uint8 tensor holding ceil(count / piece_rows) of them back to back (piece_rows = count: one piece).  They are
        unpacked into the matrix given to load_device_begin, then digested."""
        _lib.check(self._L.selb200_load_device_rows_packed(self._h, int(g0), int(count), pieces.data_ptr(),
                                                           int(piece_rows or count)))

    def load_end(self):
        _lib.check(self._L.selb200_load_end(self._h))
        self._order = None
        return self

    def load_stream(self, n: int, p: int, fill, aux_kind: int = AUX_NONE, aux_len: int = 0):
        """Streaming load over the library's pinned staging slots.  `fill(g0, count, regs, stored, aux)`
        is called per chunk with numpy views to fill in place: regs uint8[count][2^p], stored
        float64[count] (header value_, < 0 = recompute), aux uint64[count][m] / uint8[count][2^p_aux] / None."""
        rpc = C.c_int64()
        _lib.check(self._L.selb200_load_begin(self._h, n, p, aux_kind, aux_len, C.byref(rpc)))
        m = 1 << p
        for g0 in range(0, n, rpc.value):
            cnt = min(rpc.value, n - g0)
            r, st, ax = C.c_void_p(), C.c_void_p(), C.c_void_p()
            _lib.check(self._L.selb200_load_acquire(self._h, g0, cnt, C.byref(r), C.byref(st), C.byref(ax)))
            regs = np.ctypeslib.as_array(C.cast(r, C.POINTER(C.c_uint8)), shape=(cnt, m))
            stored = np.ctypeslib.as_array(C.cast(st, C.POINTER(C.c_double)), shape=(cnt,))
            aux = None
            if aux_kind == AUX_SMH:
                aux = np.ctypeslib.as_array(C.cast(ax, C.POINTER(C.c_uint64)), shape=(cnt, aux_len))
            elif aux_kind == AUX_HLL:
                aux = np.ctypeslib.as_array(C.cast(ax, C.POINTER(C.c_uint8)), shape=(cnt, 1 << aux_len))
            fill(g0, cnt, regs, stored, aux)
            _lib.check(self._L.selb200_load_commit(self._h))
        _lib.check(self._L.selb200_load_end(self._h))
        self.n, self.p, self._order, self._keep = n, p, None, None
        return self

    def order(self):
        if getattr(self, "_order", None) is not None:
            return self._order
        cards = np.empty(self.n, np.float64)
        order = np.empty(self.n, np.int32)
        _lib.check(self._L.selb200_get_order(self._h, cards.ctypes.data, order.ctypes.data))
        self._order = (cards, order)
        return self._order

    # -- multi-GPU gather over peer memory ----------------------------------------------------
    def gather_create(self, cap_pairs: int = 1 << 22) -> bytes:
        """Root rank: allocate the landing zone for the merged pair list, return the handle to hand to
        every rank's gather_attach (see include/selb200.h, selb200_gather_*)."""
        buf = C.create_string_buffer(_lib.GATHER_HANDLE_BYTES)
        _lib.check(self._L.selb200_gather_create(self._h, int(cap_pairs), buf))
        return bytes(buf.raw)

    def gather_attach(self, rank: int, world: int, handle: bytes):
        if len(handle) != _lib.GATHER_HANDLE_BYTES:
            raise ValueError("not a gather handle")
        _lib.check(self._L.selb200_gather_attach(self._h, int(rank), int(world), C.c_char_p(handle)))
        self._gather = (rank, world)
        return self

    def gather_close(self):
        self._L.selb200_gather_close(self._h)
        self._gather = None

    # -- run ------------------------------------------------------------------------------
    def run(self, tau: float = 0.9, criterion: str | int = "smh_a", z_score: float = 1.96, order_n: int = 1,
            n_rows: int = 0, n_bands: int = 0, shard: int = 0, n_shards: int = 1, sort_output: bool = True,
            fetch: bool = True, no_cb: bool = False, gather: bool = False,
            host_results: bool = False) -> SelectionResult:
        """gather=True (after gather_attach): shard/n_shards are the attached rank/world, every rank
        pushes its pairs into the root GPU's memory and the root's result is the whole job's list."""
        if gather:
            if getattr(self, "_gather", None) is None:
                raise RuntimeError("run(gather=True) needs gather_attach first")
            shard, n_shards = self._gather
        prm = _lib.Params()
        self._L.selb200_default_params(C.byref(prm))
        prm.tau = tau
        prm.criterion = CRITERIA[criterion] if isinstance(criterion, str) else int(criterion)
        prm.z_score = z_score
        prm.order_n = order_n
        prm.n_rows, prm.n_bands = n_rows, n_bands
        prm.shard, prm.n_shards = shard, n_shards
        prm.sort_output = int(sort_output)
        prm.no_cb = int(no_cb)
        prm.gather = int(gather)
        prm.host_results = int(fetch or host_results)
        st = _lib.Stats()
        _lib.check(self._L.selb200_run(self._h, C.byref(prm), C.byref(st)))
        cards, order = self.order()
        if not fetch:
            z = np.zeros(0, np.int32)
            return SelectionResult(z, z, np.zeros(0), order, cards, st.as_dict())
        cnt = self._L.selb200_result_count(self._h)
        i = np.empty(cnt, np.int32); k = np.empty(cnt, np.int32); j = np.empty(cnt, np.float64)
        _lib.check(self._L.selb200_copy_results(self._h, cnt, i.ctypes.data, k.ctypes.data, j.ctypes.data))
        nc = self._L.selb200_near_count(self._h)
        ni = np.empty(nc, np.int32); nk = np.empty(nc, np.int32); nj = np.empty(nc, np.float64)
        _lib.check(self._L.selb200_copy_near(self._h, nc, ni.ctypes.data, nk.ctypes.data, nj.ctypes.data))
        return SelectionResult(i, k, j, order, cards, st.as_dict(), ni, nk, nj)

    def result_host(self):
        """(keys uint64, jaccard float64) numpy views of the library's pinned host copy of the last run's
        lists (run(host_results=True)); valid until the next run."""
        keys, jac = C.c_void_p(), C.c_void_p()
        _lib.check(self._L.selb200_result_host(self._h, C.byref(keys), C.byref(jac)))
        cnt = self._L.selb200_result_count(self._h)
        if cnt == 0:
            return np.zeros(0, np.uint64), np.zeros(0, np.float64)
        k = np.ctypeslib.as_array(C.cast(keys, C.POINTER(C.c_uint64)), shape=(cnt,))
        j = np.ctypeslib.as_array(C.cast(jac, C.POINTER(C.c_double)), shape=(cnt,))
        return k, j

    def result_device_ptrs(self):
        keys, jac = C.c_void_p(), C.c_void_p()
        _lib.check(self._L.selb200_result_device(self._h, C.byref(keys), C.byref(jac)))
        return keys.value or 0, jac.value or 0, self._L.selb200_result_count(self._h)

    def debug_union(self, a, b) -> np.ndarray:
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32)
        t = np.empty(a.size, np.float64)
        _lib.check(self._L.selb200_debug_union(self._h, 0, a.size, a.ctypes.data, b.ctypes.data, t.ctypes.data))
        return t


def format_lines(names: list[str], res: SelectionResult) -> list[str]:
    """`fn1 + " " + fn2 + " " + std::to_string(jacc14)` (src/selection.cpp:288): %f, six decimals."""
    fi, fk = res.file_indices()
    return [f"{names[a]} {names[b]} {j:f}" for a, b, j in zip(fi.tolist(), fk.tolist(), res.jaccard.tolist())]


def run_filelist(list_file: str, tau: float = 0.9, aux_bytes: int = 256, criterion: str = "smh_a",
                 threads: int = 8, device: int = 0, base: str = "") -> list[str]:
    """The whole CLI flow of src/selection.cpp for one file list; returns the stdout lines."""
    if criterion not in CRITERIA:
        # selection.cpp:292-294
        return ["Option -c invalid. The accepted criteria are hll_a, hll_an and smh_a."]
    files = sketch_io.load_file_list(list_file)
    p, regs, stored, aux_kind, aux_len, aux = sketch_io.load_sketches(files, criterion, aux_bytes, threads, base)
    with Selection(device) as sel:
        sel.load(regs, aux, aux_kind, aux_len, stored=stored, p=p)
        res = sel.run(tau=np.float32(tau), criterion=criterion)
    return format_lines(files, res)
