"""Host mirror of the reference's `build_sketch` (src/build_sketch.cpp:186-295) over the CUDA builder.

FASTA (optionally gzip) -> per genome the record sequences joined by one 'N' -> selb200_sketch_host
-> `P.hll` plus `P.smh<m>` (-c smh_a) or `P.hll_<p>` (-c hll_a / hll_an), written in the reference's
on-disk format (sketch_io).  The sketches are byte-identical to the reference's after gunzip.
"""
from __future__ import annotations

import ctypes as C
import gzip
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib, sketch_io
from .selection import AUX_HLL, AUX_NONE, AUX_SMH


_IUPAC = set(b"ACGTURYSWKMBDHVNacgturyswkmbdhvn")
_SPACE = b" \t\n\r\v\f"


def read_fasta_clean(path: str) -> bytes:
    """Record sequences joined by one 'N', read with SeqAn's rules (seqan/seq_io/fasta_fastq.h:262-282):
    skip to '>', id = rest of the line, sequence = everything up to the next '>' with whitespace dropped;
    a character outside the IUPAC alphabet is a ParseError at which the reference stops reading the file
    (src/build_sketch.cpp:55-58), dropping that record and every later one."""
    opener = gzip.open if path.endswith(".gz") else open
    try:
        with opener(path, "rb") as f:
            raw = f.read()
    except FileNotFoundError:
        return b""                                    # "ERROR: Could not open the file": the sketch stays empty
    out = []
    for rec in raw.split(b">")[1:]:
        nl = rec.find(b"\n")
        body = rec[nl + 1:] if nl >= 0 else b""
        seq = body.translate(None, _SPACE)
        if not set(seq) <= _IUPAC:
            break
        out.append(seq)
    return b"N".join(out)


def smh_size(m_arg: int) -> int:
    return _lib.lib().selb200_smh_size(int(m_arg))


def sketch_sequences(seqs: list[bytes], p: int = 14, aux_kind: int = AUX_NONE, aux_len: int = 0, device: int = 0):
    """-> (hll uint8[n][2^p], aux array or None) for cleaned sequences."""
    L = _lib.lib()
    n = len(seqs)
    offsets = np.zeros(n + 1, np.int64)
    for i, s in enumerate(seqs):
        offsets[i + 1] = offsets[i] + len(s)
    blob = np.frombuffer(b"".join(seqs), dtype=np.uint8) if offsets[-1] else np.zeros(1, np.uint8)
    blob = np.ascontiguousarray(blob)
    hll = np.empty((n, 1 << p), np.uint8)
    aux = None
    if aux_kind == AUX_SMH:
        aux = np.empty((n, smh_size(aux_len)), np.uint64)
    elif aux_kind == AUX_HLL:
        aux = np.empty((n, 1 << aux_len), np.uint8)
    rc = L.selb200_sketch_host(device, n, blob.ctypes.data, offsets.ctypes.data, p, aux_kind, aux_len,
                               hll.ctypes.data, aux.ctypes.data if aux is not None else None)
    if rc != 0:
        raise _lib.SelB200Error(rc, L.selb200_sketch_last_error().decode("utf-8", "replace"))
    return hll, aux


def build_filelist(list_file: str, aux_bytes: int = 256, criterion: str = "", threads: int = 8, device: int = 0,
                   base: str = "", out_base: str | None = None) -> int:
    """`build_sketch -l list -t threads -a aux_bytes -c criterion`.  Returns the number of genomes."""
    import os
    files = sketch_io.load_file_list(list_file)
    aux_kind, aux_len = AUX_NONE, 0
    if criterion == "smh_a":
        aux_kind, aux_len = AUX_SMH, aux_bytes // 8                         # build_sketch.cpp:274
    elif criterion in ("hll_a", "hll_an"):
        aux_kind, aux_len = AUX_HLL, (aux_bytes & -aux_bytes).bit_length() - 1   # build_sketch.cpp:243,259
    src = [os.path.join(base, f) if base else f for f in files]
    dst = [os.path.join(out_base, f) if out_base is not None else s for f, s in zip(files, src)]
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        seqs = list(ex.map(read_fasta_clean, src))
    hll, aux = sketch_sequences(seqs, 14, aux_kind, aux_len, device)

    def write(i):
        os.makedirs(os.path.dirname(dst[i]) or ".", exist_ok=True)
        sketch_io.write_hll(dst[i] + ".hll", hll[i], 14)                    # build_sketch.cpp:237
        if aux_kind == AUX_SMH:
            sketch_io.write_smh(dst[i] + ".smh" + str(aux_bytes // 8), aux[i])   # :288 (name uses a/8 unrounded)
        elif aux_kind == AUX_HLL:
            sketch_io.write_hll(dst[i] + ".hll_" + str(aux_len), aux[i], aux_len)

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(write, range(len(files))))
    if criterion not in ("smh_a", "hll_a", "hll_an"):
        print("Option -c invalid. The accepted criteria are hll_a, hll_an and smh_a.")   # build_sketch.cpp:290-292
    return len(files)
