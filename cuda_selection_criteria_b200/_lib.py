"""ctypes binding of include/selb200.h.  Loading fails loudly when the library is not built."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libselb200.so"


class LibraryNotBuilt(RuntimeError):
    pass


class SelB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"selb200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("tau", C.c_float), ("criterion", C.c_int32), ("z_score", C.c_float), ("order_n", C.c_int32),
        ("n_rows", C.c_int32), ("n_bands", C.c_int32), ("shard", C.c_int32), ("n_shards", C.c_int32),
        ("sort_output", C.c_int32), ("no_cb", C.c_int32), ("gather", C.c_int32), ("host_results", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("pairs_total", C.c_int64), ("pairs_cb", C.c_int64), ("pairs_cb_shard", C.c_int64),
        ("pairs_cand", C.c_int64), ("pairs_aux", C.c_int64), ("pairs_out", C.c_int64), ("pairs_near", C.c_int64),
        ("tiles_total", C.c_int64), ("tiles_shard", C.c_int64),
        ("n_bands", C.c_int32), ("n_rows", C.c_int32), ("batches", C.c_int32), ("launches", C.c_int32),
        ("ms_bounds", C.c_float), ("ms_filter", C.c_float), ("ms_verify", C.c_float), ("ms_union", C.c_float),
        ("ms_estimate", C.c_float), ("ms_sort", C.c_float), ("ms_total", C.c_float), ("reserved0", C.c_int32),
        ("filter_steps", C.c_int64), ("reserved", C.c_int32 * 5),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


GATHER_HANDLE_BYTES = 128

# every symbol include/selb200.h declares: (name, restype, argtypes)
_VP, _I, _I64 = C.c_void_p, C.c_int, C.c_int64
SYMBOLS = [
    ("selb200_abi_version", _I, []),
    ("selb200_last_error", C.c_char_p, []),
    ("selb200_device_count", _I, []),
    ("selb200_create", _I, [_I, _VP, C.POINTER(_VP)]),
    ("selb200_destroy", None, [_VP]),
    ("selb200_load_host", _I, [_VP, _I64, _I, _VP, _VP, _I, _I, _VP]),
    ("selb200_load_device", _I, [_VP, _I64, _I, _VP, _VP, _I, _I, _VP]),
    ("selb200_load_device_begin", _I, [_VP, _I64, _I, _VP, _I, _I, _VP]),
    ("selb200_load_device_rows", _I, [_VP, _I64, _I64]),
    ("selb200_load_info", _I, [_VP, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    ("selb200_nib4_piece_bytes", _I64, [_I64, _I]),
    ("selb200_nib4_pack_piece", _I64, [_VP, _I64, _I, _VP, _I]),
    ("selb200_load_device_rows_packed", _I, [_VP, _I64, _I64, _VP, _I64]),
    ("selb200_load_begin", _I, [_VP, _I64, _I, _I, _I, C.POINTER(_I64)]),
    ("selb200_load_acquire", _I, [_VP, _I64, _I64, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP)]),
    ("selb200_load_commit", _I, [_VP]),
    ("selb200_load_end", _I, [_VP]),
    ("selb200_get_order", _I, [_VP, _VP, _VP]),
    ("selb200_default_params", None, [C.POINTER(Params)]),
    ("selb200_run", _I, [_VP, C.POINTER(Params), C.POINTER(Stats)]),
    ("selb200_result_count", _I64, [_VP]),
    ("selb200_copy_results", _I, [_VP, _I64, _VP, _VP, _VP]),
    ("selb200_near_count", _I64, [_VP]),
    ("selb200_copy_near", _I, [_VP, _I64, _VP, _VP, _VP]),
    ("selb200_result_device", _I, [_VP, C.POINTER(_VP), C.POINTER(_VP)]),
    ("selb200_result_host", _I, [_VP, C.POINTER(_VP), C.POINTER(_VP)]),
    ("selb200_gather_create", _I, [_VP, _I64, _VP]),
    ("selb200_gather_attach", _I, [_VP, _I, _I, _VP]),
    ("selb200_gather_close", None, [_VP]),
    ("selb200_band_params", _I, [_I, C.c_float, _I, C.POINTER(_I), C.POINTER(_I)]),
    ("selb200_sort_order", _I, [_I64, _VP, _VP]),
    ("selb200_debug_union", _I, [_VP, _I, _I64, _VP, _VP, _VP]),
    ("selb200_smh_size", _I, [_I]),
    ("selb200_sketch_host", _I, [_I, _I64, _VP, _VP, _I, _I, _I, _VP, _VP]),
    ("selb200_sketch_last_error", C.c_char_p, []),
    ("selb200_warmup", _I, [_I]),
    ("selb200_synth_hll", _I, [_I, _I, _I64, _I, _VP, _I64, _VP, _VP, C.c_uint64, C.c_uint32, _VP]),
    ("selb200_synth_smh", _I, [_I, _I, _I64, _I, _VP, _I64, _VP, _VP, C.c_uint64, C.c_uint32, _VP]),
]

_lib = None


def lib_path() -> str:
    # SELB200_LIB: another build of the same library (kernel-variant A/B measurements)
    return os.environ.get("SELB200_LIB") or os.path.join(_HERE, _LIB_NAME)


def lib() -> C.CDLL:
    """The loaded C-ABI library.  Raises LibraryNotBuilt if csrc/ has not been compiled."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise LibraryNotBuilt(
            f"{path} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    L = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(L, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if L.selb200_abi_version() != 1:
        raise LibraryNotBuilt(f"{path}: ABI version {L.selb200_abi_version()} != 1, rebuild")
    _lib = L
    return L


def check(code: int) -> None:
    if code != 0:
        raise SelB200Error(code, lib().selb200_last_error().decode("utf-8", "replace"))
