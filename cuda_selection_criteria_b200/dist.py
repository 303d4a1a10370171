"""Multi-GPU sharding of the pair triangle: one process per GPU, torch.distributed for the plumbing.

The path shards by construction (pairs are independent): every rank holds the full sorted
sketch matrices (1.64 GB + 0.1 GB at n=100k; broadcast from rank 0 over NCCL/NVLink, or all-gathered
from per-rank host slices: ShardedSketches), runs tiles r, r+R, r+2R, ... of the CB-band tile list,
and the variable-length pair lists meet on rank 0.  The product path for that last step is the
library's own peer-memory gather (setup_gather + Selection.run(gather=True): kernels store into
the root GPU's memory over NVLink, no collective call); gather_lists is the NCCL/gloo
restatement of the same exchange, kept as its checker and for CPU tests.  There is no collective
inside the compare kernels — the only exchange steps are the sketches in and the lists out
(SURVEY.md §8e).  Works with backend "nccl" (CUDA tensors)
and "gloo" (CPU tensors; used by the world_size-2 CPU tests of this plumbing).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class _DevView:
    """Zero-copy view of a raw device pointer for torch.as_tensor."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, n: int, typestr: str, device: int) -> torch.Tensor:
    if n == 0 or not ptr:
        dt = {"<i8": torch.int64, "<f8": torch.float64}[typestr]
        return torch.empty(0, dtype=dt, device=f"cuda:{device}")
    return torch.as_tensor(_DevView(ptr, n, typestr), device=f"cuda:{device}")


def shard_tiles(tiles_total: int, rank: int, world: int) -> range:
    """Tiles of one shard: the row-major tile list dealt round-robin (same rule as selb200_run)."""
    return range(rank, tiles_total, world)


def broadcast_sketches(regs: torch.Tensor, aux: torch.Tensor | None, src: int = 0):
    """In-place broadcast of the sketch matrices from `src` (every rank passes same-shape buffers)."""
    dist.broadcast(regs, src=src)
    if aux is not None:
        dist.broadcast(aux, src=src)


def slice_rows(n: int, rank: int, world: int) -> tuple[int, int, int]:
    """(first row, row count, rows per slice) of `rank`'s contiguous slice of n genomes in file-list order."""
    per = (n + world - 1) // world
    g0 = min(n, rank * per)
    return g0, min(n, g0 + per) - g0, per


class ShardedSketches:
    """Device-resident sketch matrices assembled from per-rank HOST slices.

    Every rank owns rows [r*per, (r+1)*per) of the file list in (pinned) host memory — the natural
    state after each rank has decoded its share of the sketch files.  `assemble` copies the slice to
    the rank's GPU over its own PCIe link and all-gathers the slices over NCCL/NVLink, so the
    host->device leg of an R-rank job runs on R links at once instead of through rank 0
    (broadcast_sketches).  The slice travels in `chunks` pieces: piece c+1 is still on the PCIe bus
    (side stream) while piece c is being all-gathered, in place — the device matrix is laid out
    piece-major, [piece][rank][rows], so that each all-gather writes one contiguous block.
    Device row (c*R + r)*rpc + j therefore holds file `row_to_file[row]` = r*per + c*rpc + j; rows past
    a rank's share are all-zero sketches (cardinality 0: they sort first and pair with nothing)."""

    def __init__(self, n: int, m: int, aux_cols: int, aux_dtype, device, rank: int, world: int, chunks: int = 4):
        self.n, self.rank, self.world = n, rank, world
        self.g0, self.rows, self.per = slice_rows(n, rank, world)
        self.chunks = max(1, min(chunks, self.per))
        self.rpc = (self.per + self.chunks - 1) // self.chunks          # rows per piece
        self.n_dev = self.chunks * world * self.rpc
        self.regs = torch.zeros((self.n_dev, m), dtype=torch.uint8, device=device)
        self.aux = torch.zeros((self.n_dev, aux_cols), dtype=aux_dtype, device=device) if aux_cols else None
        c, r, j = np.meshgrid(np.arange(self.chunks), np.arange(world), np.arange(self.rpc), indexing="ij")
        local = c * self.rpc + j                                         # row inside the rank's slice
        f = r * self.per + local
        ok = (local < self.per) & (f < n)
        self.row_to_file = np.where(ok, f, -1).reshape(-1).astype(np.int64)
        self._copy_stream = torch.cuda.Stream(device=device) if torch.device(device).type == "cuda" else None

    def _piece(self, t: torch.Tensor, c: int):
        lo = c * self.world * self.rpc
        return t[lo:lo + self.world * self.rpc], t[lo + self.rank * self.rpc:lo + (self.rank + 1) * self.rpc]

    SUB_ROWS_BYTES = 16 << 20      # registers packed at a time: four pinned staging slots of this size stay in the host's
                                   # last-level cache, so the copy engine reads the packed bytes from there, not from DRAM

    def _packed_buffers(self, m: int):
        """Pinned host staging and device landing zones of the packed rows (include/selb200.h, "Packed transport"): a
        piece (the unit of the all-gather) is packed as ceil(rpc / sub_rows) small pieces back to back, [piece][rank][...]
        on the device, so that one all-gather per piece moves every rank's rows."""
        if getattr(self, "_pk_dev", None) is None:
            L = _lib.lib()
            self._p = m.bit_length() - 1
            self._sub_rows = max(1, min(self.rpc, self.SUB_ROWS_BYTES // m))
            self._n_sub = (self.rpc + self._sub_rows - 1) // self._sub_rows
            self._sub_bytes = int(L.selb200_nib4_piece_bytes(self._sub_rows, self._p))
            if self._sub_bytes <= 0:
                raise _lib.SelB200Error(L.selb200_last_error().decode())
            self._pk_dev = torch.empty((self.chunks, self.world, self._n_sub * self._sub_bytes), dtype=torch.uint8,
                                       device=self.regs.device)
            self._pk_host = [torch.empty(self._sub_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
            self._pk_free = [None] * 4
            self._pad = torch.zeros((self._sub_rows, m), dtype=torch.uint8)
        return self._pk_dev, self._pk_host

    def _pack_and_copy(self, L, regs_host, c: int, h0: int, h1: int):
        """rows [h0, h1) of this rank's slice (piece c, rpc rows with the zero padding) -> packed, into pk_dev[c, rank];
        the copies are queued on the copy stream, the host only waits for a staging slot to come back"""
        slot = getattr(self, "_slot", 0)
        import time as _t
        tr = self._trace
        for s in range(self._n_sub):
            r0 = s * self._sub_rows
            rows = min(self._sub_rows, self.rpc - r0)
            a, b = min(h1, h0 + r0), min(h1, h0 + r0 + rows)
            if b - a == rows:
                src = regs_host[a:b]
            else:                       # the slice ends inside this piece: the rest are all-zero sketches
                self._pad.zero_()
                if b > a:
                    self._pad[:b - a].copy_(regs_host[a:b])
                src = self._pad
            t0 = _t.perf_counter()
            if self._pk_free[slot] is not None:
                self._pk_free[slot].synchronize()
            t1 = _t.perf_counter()
            n_raw = int(L.selb200_nib4_pack_piece(src.data_ptr(), rows, self._p, self._pk_host[slot].data_ptr(), 0))
            t2 = _t.perf_counter()
            tr[0] += t1 - t0
            tr[1] += t2 - t1
            if n_raw < 0:
                raise _lib.SelB200Error(L.selb200_last_error().decode())
            if n_raw > 4:
                raise _lib.SelB200Error(f"{n_raw} rows of a piece do not look like HLL sketches (more than 32 registers far "
                                        "above the smallest): set SELB200_H2D=raw to move the registers unpacked")
            nbytes = int(L.selb200_nib4_piece_bytes(rows, self._p))
            with torch.cuda.stream(self._copy_stream):
                dst = self._pk_dev[c, self.rank, s * self._sub_bytes:s * self._sub_bytes + nbytes]
                dst.copy_(self._pk_host[slot][:nbytes], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._pk_free[slot] = ev
            slot = (slot + 1) & 3
        self._slot = slot

    def assemble(self, regs_host: torch.Tensor, aux_host: torch.Tensor | None, on_piece=None, on_piece_packed=None):
        """on_piece(row0, rows): called after the all-gather of each piece has been queued on the current
        stream (e.g. Selection.load_device_rows, so that the piece is digested while the next one travels).

        With on_piece_packed (e.g. Selection.load_device_rows_packed) the register rows travel PACKED — half the
        bytes over PCIe and over NVLink: every piece is packed by this rank's host threads into pinned memory, copied,
        all-gathered as bytes, and on_piece_packed(row0, rows, pieces, piece_rows) is called for each rank's part, which
        unpacks it into `self.regs` and digests it.  Default for a world of one; SELB200_H2D=packed selects it for any world."""
        import time as _t
        t_begin = _t.perf_counter()
        t_cb = 0.0
        cuda = self._copy_stream is not None
        # packing costs host memory bandwidth (about 110 GB/s of registers on a 24-thread box, all ranks together), a raw
        # copy costs PCIe time (54 GB/s per link): packed wins on one link (18 ms against 32 for 1.64 GB), loses from
        # two links on (22 ms against 18.5 at two ranks), so it is the default for a world of one only;
        # SELB200_H2D=packed / raw force either route
        mode = os.environ.get("SELB200_H2D", "")
        packed = cuda and on_piece_packed is not None and (mode == "packed" or (mode != "raw" and self.world == 1))
        if cuda:
            self._copy_stream.wait_stream(torch.cuda.current_stream())   # earlier readers of the matrices are done
        if packed:
            pk_dev, _ = self._packed_buffers(int(regs_host.shape[1]))
            L = _lib.lib()
            self._trace = [0.0, 0.0]
        for c in range(self.chunks):
            h0, h1 = min(self.rows, c * self.rpc), min(self.rows, (c + 1) * self.rpc)
            if packed:
                self._pack_and_copy(L, regs_host, c, h0, h1)
            ctx = torch.cuda.stream(self._copy_stream) if cuda else _Null()
            with ctx:
                if not packed and h1 > h0:
                    self._piece(self.regs, c)[1][:h1 - h0].copy_(regs_host[h0:h1], non_blocking=True)
                if h1 > h0 and self.aux is not None:
                    self._piece(self.aux, c)[1][:h1 - h0].copy_(aux_host[h0:h1], non_blocking=True)
                if cuda:
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
            # queued behind the copies on the current stream: the host goes on to pack the next piece meanwhile
            if cuda:
                torch.cuda.current_stream().wait_event(ev)
            if self.world > 1:
                if packed:
                    dist.all_gather_into_tensor(pk_dev[c].view(-1), pk_dev[c, self.rank])
                else:
                    whole, mine = self._piece(self.regs, c)
                    dist.all_gather_into_tensor(whole, mine)
                if self.aux is not None:
                    whole, mine = self._piece(self.aux, c)
                    dist.all_gather_into_tensor(whole, mine)
            if packed:
                t0 = _t.perf_counter()
                for r in range(self.world):
                    on_piece_packed((c * self.world + r) * self.rpc, self.rpc, pk_dev[c, r], self._sub_rows)
                t_cb += _t.perf_counter() - t0
            elif on_piece is not None:
                on_piece(c * self.world * self.rpc, self.world * self.rpc)
        if packed and os.environ.get("SELB200_PACK_TRACE"):
            import sys
            print(f"rank {self.rank} packed assemble: waiting for a slot {self._trace[0] * 1e3:.2f} ms, pack {self._trace[1] * 1e3:.2f} ms, "
                  f"callbacks {t_cb * 1e3:.2f} ms, whole call {(_t.perf_counter() - t_begin) * 1e3:.2f} ms", file=sys.stderr)
        return self.regs, self.aux


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def setup_gather(sel, cap_pairs: int = 1 << 22, root: int = 0):
    """Create the root's landing zone and attach every rank's context to it (Selection.gather_*).
    The handle travels through the process group's object broadcast (any backend)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [sel.gather_create(cap_pairs) if rank == root else None]
    dist.broadcast_object_list(box, src=root)
    sel.gather_attach(rank, world, box[0])
    dist.barrier()
    return sel


def gather_lists(keys: torch.Tensor, jac: torch.Tensor, dst: int = 0):
    """Gather per-rank (keys int64, jaccard float64) lists on `dst`, merged in (i,k) order.

    One all_gather of the counts (a single host sync) and one all_gather of a padded
    [keys | jaccard-bits] payload.  Returns (keys, jaccard) on dst, (None, None) elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n = keys.numel()
    cnt = torch.tensor([n], dtype=torch.int64, device=keys.device)
    counts_t = torch.empty(world, dtype=torch.int64, device=keys.device)
    dist.all_gather_into_tensor(counts_t, cnt)
    counts = counts_t.tolist()
    mx = max(counts) if counts else 0
    if mx == 0:
        return (keys[:0], jac[:0]) if rank == dst else (None, None)
    buf = torch.empty(2 * mx, dtype=torch.int64, device=keys.device)
    buf[:n] = keys
    buf[mx:mx + n] = jac.view(torch.int64)
    out = torch.empty(world * 2 * mx, dtype=torch.int64, device=keys.device)
    dist.all_gather_into_tensor(out, buf)
    if rank != dst:
        return None, None
    out = out.view(world, 2, mx)
    allk = torch.cat([out[r, 0, :c] for r, c in enumerate(counts)])
    allj = torch.cat([out[r, 1, :c] for r, c in enumerate(counts)]).view(torch.float64)
    # a device sort of the merged list restores the reference's print order exactly
    o = torch.argsort(allk, stable=True)
    return allk[o], allj[o]


def split_keys(keys: torch.Tensor):
    k = keys.cpu().numpy()
    return (k >> 32).astype(np.int32), (k & 0xFFFFFFFF).astype(np.int32)
