"""On-disk sketch formats of the reference, read and written byte-exactly (host side).

* ``P.hll`` / ``P.hll_<p>`` — sketch::hll_t::write/read (sketch/include/sketch/hll.h:1103-1111,
  1126-1143): gzip stream of ``u32[4] = {is_calculated, estim, jestim, 1}``, ``u32 np``,
  ``f64 value`` (-1.0 if not computed), ``u8 core[2^np]``.
* ``P.smh<m>`` — write_smh / read_smh (src/build_sketch.cpp:9-20, src/selection.cpp:12-33):
  gzip stream of ``u32 m`` then ``u64 h[m]``.
File-name rules: src/selection.cpp:125,138-139,231,245-246.
"""
from __future__ import annotations

import gzip
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ERTL_MLE = 2
_HLL_HDR = struct.Struct("<5Id")


class SketchFormatError(RuntimeError):
    pass


def read_hll(path: str):
    """-> (p, estim, jestim, stored_value, registers uint8[2^p])."""
    try:
        with gzip.open(path, "rb") as f:
            data = f.read()
    except FileNotFoundError:
        # selection.cpp:16 / hll.h:1146: "Could not open file at '<path>' for reading"
        raise FileNotFoundError(f"Could not open file at '{path}' for reading") from None
    if len(data) < _HLL_HDR.size:
        raise SketchFormatError(f"Error reading from file {path}")
    _calc, estim, jestim, _one, np_, value = _HLL_HDR.unpack_from(data)
    m = 1 << np_
    if len(data) < _HLL_HDR.size + m:
        raise SketchFormatError(f"Error reading from file {path}")
    regs = np.frombuffer(data, dtype=np.uint8, count=m, offset=_HLL_HDR.size)
    return np_, estim, jestim, value, regs


def write_hll(path: str, regs: np.ndarray, p: int, value: float = -1.0, level: int = 6) -> None:
    hdr = _HLL_HDR.pack(1 if value >= 0 else 0, ERTL_MLE, ERTL_MLE, 1, p, value)
    with gzip.open(path, "wb", compresslevel=level) as f:
        f.write(hdr)
        f.write(np.ascontiguousarray(regs, dtype=np.uint8).tobytes())


def read_smh(path: str) -> np.ndarray:
    try:
        with gzip.open(path, "rb") as f:
            data = f.read()
    except FileNotFoundError:
        raise FileNotFoundError(f"Could not open file at '{path}' for reading") from None
    if len(data) < 4:
        raise SketchFormatError("Error reading from file\n")
    (m,) = struct.unpack_from("<I", data)
    if len(data) < 4 + 8 * m:
        raise SketchFormatError("Error reading from file\n")
    return np.frombuffer(data, dtype=np.uint64, count=m, offset=4)


def write_smh(path: str, buckets: np.ndarray, level: int = 6) -> None:
    b = np.ascontiguousarray(buckets, dtype=np.uint64)
    with gzip.open(path, "wb", compresslevel=level) as f:
        f.write(struct.pack("<I", b.size))
        f.write(b.tobytes())


def load_file_list(list_file: str, prefix: str = "") -> list[str]:
    """src/selection.cpp:36-63: one path per line, surrounding ' \\t\\r\\n' trimmed, empty lines skipped."""
    if not list_file:
        raise SystemExit("No input file provided")
    try:
        fh = open(list_file, "r")
    except OSError:
        raise SystemExit("No valid input file provided") from None
    files = []
    with fh:
        for line in fh:
            line = line.strip(" \t\r\n")
            if line:
                files.append(prefix + line)
    return files


def aux_suffix(criterion: str, aux_bytes: int) -> str:
    if criterion == "smh_a":
        return ".smh" + str(aux_bytes // 8)                      # selection.cpp:231,246
    p = (aux_bytes & -aux_bytes).bit_length() - 1                 # __builtin_ctz, selection.cpp:125
    return ".hll_" + str(p)                                       # selection.cpp:139


def load_sketches(files: list[str], criterion: str, aux_bytes: int, threads: int = 8, base: str = ""):
    """Parallel gunzip of every sketch the reference's loader opens (selection.cpp:134-142,241-249).

    Returns (p, regs uint8[n][2^p], stored float64[n], aux_kind, aux_len, aux array or None).
    """
    import os

    n = len(files)
    suffix = aux_suffix(criterion, aux_bytes) if criterion in ("smh_a", "hll_a", "hll_an") else None

    def one(i):
        path = os.path.join(base, files[i]) if base else files[i]
        h = read_hll(path + ".hll")
        a = None
        if suffix is not None:
            a = read_smh(path + suffix) if criterion == "smh_a" else read_hll(path + suffix)
        return h, a

    if n == 0:
        return 14, np.zeros((0, 1 << 14), np.uint8), np.zeros(0), 0, 0, None
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        out = list(ex.map(one, range(n)))
    p = out[0][0][0]
    regs = np.empty((n, 1 << p), dtype=np.uint8)
    stored = np.empty(n, dtype=np.float64)
    for i, (h, _a) in enumerate(out):
        if h[0] != p:
            raise SketchFormatError(f"{files[i]}.hll has p={h[0]}, expected {p}")
        if h[1] != ERTL_MLE or h[2] != ERTL_MLE:
            raise SketchFormatError(
                f"{files[i]}.hll stores estimator ({h[1]},{h[2]}); only ERTL_MLE (2,2), the build_sketch "
                "default (hll.h:825-827), is supported")
        regs[i] = h[4]
        stored[i] = h[3]
    if criterion == "smh_a":
        m = out[0][1].size
        aux = np.empty((n, m), dtype=np.uint64)
        for i, (_h, a) in enumerate(out):
            if a.size != m:
                raise SketchFormatError(f"{files[i]}{suffix}: {a.size} buckets, expected {m}")
            aux[i] = a
        return p, regs, stored, 1, m, aux
    if criterion in ("hll_a", "hll_an"):
        pa = out[0][1][0]
        aux = np.empty((n, 1 << pa), dtype=np.uint8)
        for i, (_h, a) in enumerate(out):
            if a[0] != pa:
                raise SketchFormatError(f"{files[i]}{suffix} has p={a[0]}, expected {pa}")
            aux[i] = a[4]
        return p, regs, stored, 2, pa, aux
    return p, regs, stored, 0, 0, None


def write_dataset(dirpath: str, names: list[str], p: int, regs: np.ndarray, smh: np.ndarray | None = None,
                  aux_hll: np.ndarray | None = None, aux_p: int = 0, threads: int = 8, level: int = 1) -> str:
    """Write .hll (+ .smh<m> / .hll_<p>) files and a file list the unmodified reference can consume."""
    import os

    os.makedirs(dirpath, exist_ok=True)

    def one(i):
        base = os.path.join(dirpath, names[i])
        write_hll(base + ".hll", regs[i], p, level=level)
        if smh is not None:
            write_smh(base + ".smh" + str(smh.shape[1]), smh[i], level=level)
        if aux_hll is not None:
            write_hll(base + ".hll_" + str(aux_p), aux_hll[i], aux_p, level=level)

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(one, range(len(names))))
    lst = os.path.join(dirpath, "filelist.txt")
    with open(lst, "w") as f:
        for nm in names:
            f.write(os.path.join(dirpath, nm) + "\n")
    return lst


_ = zlib  # gzip's C implementation releases the GIL; threads scale
