#!/usr/bin/env python
"""bench.py — sketch-pair comparisons/sec of the all-pairs selection path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's OpenMP selection.cpp

Workload (config C4 of SURVEY.md §8d, the configuration the metric is quoted on): n = 100 000
synthetic bacterial-size genomes ("synth-v1", seed 1002), primary HLL p=14, CB + smh_a with a
1 KiB SuperMinHash auxiliary sketch (m=128), tau = 0.9.  A *step* is one pass of the hot path
over that batch: CB band -> smh_a filter -> HLL-14 union -> Jaccard >= tau -> pair list, sorted,
gathered on rank 0.  metric = [n(n-1)/2] / step time.

  value : inputs already resident in HBM (sketches generated on the device; with N>1 ranks they
          are broadcast once over NCCL before the timed region), device time by CUDA events on
          the stream the kernels run on, max over ranks.
  e2e   : the same through the public API with HOST buffers: pinned host -> device copy of the
          sketch matrices, per-genome cardinalities + sort, the run, and the device -> host read
          of the pair list, every step (N>1: rank 0 uploads, NCCL broadcast, shards, gather).
  roofline / cpu_baseline: see DESIGN.md §measurement.

One JSON line on stdout (rank 0).  Under torchrun it reads RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sketch_pair_comparisons_per_sec"
UNIT = "pairs/s"
ALG_BYTES = {"cb": 16, "smh_a": 2048, "union": 32768, "emit": 16}   # SURVEY.md §8d, per unit of work


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--genomes", dest="n", type=int, default=100_000,
                    help="genomes (default: config C4); use --genomes under torchrun, whose own parser claims --n")
    ap.add_argument("--criterion", default="smh_a", choices=["cb", "smh_a", "hll_a", "hll_an"])
    ap.add_argument("--tau", type=float, default=0.9)
    ap.add_argument("--aux-bytes", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=1002)
    ap.add_argument("--cpu-sample", type=int, default=25_000, help="genomes in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"synth-v1 seed {a.seed}: n={a.n} genomes, HLL p=14, CB + {a.criterion}"
            + (f" (aux {a.aux_bytes} B)" if a.criterion != "cb" else "") + f", tau={a.tau}")


# ----------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu), "-lms", "100"], stdout=open(self.path, "w"),
                                     stderr=subprocess.DEVNULL)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(kernel: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t.get(kernel)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# CPU baseline: the reference's own OpenMP selection.cpp (oracle/_ref/selection), bounded sample
# ----------------------------------------------------------------------------------------------
def write_sample(a, n_s: int, td: str):
    from cuda_selection_criteria_b200 import sketch_io, synth
    plan = synth.make_plan(a.n, a.seed).head(n_s)
    regs = synth.hll(plan, 14)
    names = [f"g{i:06d}.fna.gz" for i in range(n_s)]
    smh = aux = None
    pa = 0
    if a.criterion == "smh_a":
        smh = synth.smh(plan, a.aux_bytes // 8)
    elif a.criterion in ("hll_a", "hll_an"):
        pa = a.aux_bytes.bit_length() - 1
        aux = synth.hll(plan, pa, synth.TAG_AUX_HLL)
    else:
        smh = np.full((n_s, 1), 42, np.uint64)     # CB only: constant .smh1 -> smh_a always passes
    sketch_io.write_dataset(td, names, 14, regs, smh=smh, aux_hll=aux, aux_p=pa, threads=os.cpu_count() or 8)
    with open(os.path.join(td, "list.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    return regs, smh if smh is not None else aux


def ref_flags(a):
    crit = "smh_a" if a.criterion == "cb" else a.criterion
    aux_bytes = 8 if a.criterion == "cb" else a.aux_bytes
    return crit, aux_bytes


def time_reference_once(a, td: str, cores: int):
    """compare-phase seconds of the unmodified binary: wall(tau) - wall(-h 2.0) (load + sort only)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "selection")
    crit, aux_bytes = ref_flags(a)
    base = [ref, "-l", "list.txt", "-t", str(cores), "-a", str(aux_bytes), "-c", crit]
    t0 = time.perf_counter()
    out = subprocess.run(base + ["-h", str(a.tau)], cwd=td, capture_output=True, check=True).stdout
    t1 = time.perf_counter()
    subprocess.run(base + ["-h", "2.0"], cwd=td, capture_output=True, check=True)
    t2 = time.perf_counter()
    return max((t1 - t0) - (t2 - t1), 1e-9), (t1 - t0), (t2 - t1), out.count(b"\n")


def time_port_once(a, regs, aux, cores: int):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    t0 = time.perf_counter()
    res = O.select(regs, 14, a.criterion, np.float32(a.tau), aux=aux, threads=cores)
    # the port computes cardinalities inside; subtract nothing — report as is, it is only a fallback
    return time.perf_counter() - t0, len(res["i"])


def cpu_baseline(a, steps: int = 1, warmup: int = 0):
    cores = os.cpu_count() or 1
    n_s = min(a.cpu_sample, a.n)
    pairs = n_s * (n_s - 1) / 2
    td = tempfile.mkdtemp(prefix="selb200_cpu_")
    try:
        regs, aux = write_sample(a, n_s, td)
        have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "selection"))
        secs = []
        detail = {}
        for it in range(warmup + steps):
            if have_ref:
                cmp_s, full_s, load_s, lines = time_reference_once(a, td, cores)
                detail = {"wall_tau_s": round(full_s, 3), "wall_load_only_s": round(load_s, 3), "lines": lines}
            else:
                cmp_s, lines = time_port_once(a, regs, aux if a.criterion != "cb" else None, cores)
                detail = {"lines": lines}
            if it >= warmup:
                secs.append(cmp_s)
        t = sum(secs) / len(secs)
        return {"value": pairs / t, "unit": UNIT, "cores": cores, "kind": "reference" if have_ref else "port",
                "sample": (f"first {n_s} genomes of the workload ({int(pairs)} pairs); compare phase = "
                           f"wall(-h {a.tau}) - wall(-h 2.0) of oracle/_ref/selection -t {cores}"
                           if have_ref else f"first {n_s} genomes, oracle port incl. cardinalities, {cores} threads"),
                "seconds_per_pass": t, **detail}
    finally:
        shutil.rmtree(td, ignore_errors=True)


def run_reference_arm(a, rank: int, world: int, out):
    if rank != 0:
        return
    cb = cpu_baseline(a, steps=a.steps, warmup=a.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["seconds_per_pass"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/u64 + f64",
            "data": "synthetic", "config": {"workload": workload_name(a), "sample": cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=out, flush=True)


# ----------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------
def _protect_stdout():
    """Only the JSON line may reach stdout: libraries (NCCL's version banner, torchrun notes) write
    there too, so fd 1 is pointed at stderr for the whole run and the line goes to the saved fd."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    a = parse()
    real_stdout = _protect_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world, real_stdout)
        return

    import torch
    import torch.distributed as dist
    import cuda_selection_criteria_b200 as S
    from cuda_selection_criteria_b200 import dist as sdist, synth
    from cuda_selection_criteria_b200.selection import AUX_HLL, AUX_NONE, AUX_SMH

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    aux_kind = {"cb": AUX_NONE, "smh_a": AUX_SMH, "hll_a": AUX_HLL, "hll_an": AUX_HLL}[a.criterion]
    plan = synth.make_plan(a.n, a.seed)
    # ---- resident inputs: generated on rank 0's device, broadcast over NCCL -----------------------
    if rank == 0:
        regs_d = synth.hll(plan, 14, device=local)
        aux_d = None
        if aux_kind == AUX_SMH:
            aux_d = synth.smh(plan, a.aux_bytes // 8, device=local)
        elif aux_kind == AUX_HLL:
            aux_d = synth.hll(plan, a.aux_bytes.bit_length() - 1, synth.TAG_AUX_HLL, device=local)
    else:
        regs_d = torch.empty((a.n, 1 << 14), dtype=torch.uint8, device=dev)
        aux_d = None
        if aux_kind == AUX_SMH:
            aux_d = torch.empty((a.n, a.aux_bytes // 8), dtype=torch.int64, device=dev)
        elif aux_kind == AUX_HLL:
            aux_d = torch.empty((a.n, a.aux_bytes), dtype=torch.uint8, device=dev)
    if world > 1:
        sdist.broadcast_sketches(regs_d, aux_d, src=0)
    torch.cuda.synchronize()

    # one explicit stream for everything: torch copies and NCCL collectives issued under it and the
    # library's kernels (which run on the stream handed to the context) are then ordered on the device
    # (the legacy default stream has handle 0, which the C-ABI reads as "use a stream of your own")
    main_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(main_stream)
    stream = main_stream.cuda_stream
    assert stream != 0
    sel = S.Selection(local, stream=stream)
    sel.load(regs_d, aux_d, aux_kind)

    tau32 = np.float32(a.tau)
    stats_acc = []

    if world > 1:
        sdist.setup_gather(sel)        # rank 0's landing zone, mapped into every rank (CUDA IPC)

    def step_resident():
        # with several ranks every rank's kernels store its pairs into rank 0's memory over NVLink
        # (peer-memory gather inside selb200_run); rank 0 sorts the merged list and its run ends with the
        # D2H of that list into pinned host memory — inside the metric
        res = sel.run(tau=tau32, criterion=a.criterion, shard=rank, n_shards=world, fetch=False,
                      gather=(world > 1), host_results=(rank == 0))
        return res.stats, (sel.result_host()[0].size if rank == 0 else 0)

    for _ in range(a.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    out_n = 0
    for _ in range(a.steps):
        st, out_n = step_resident()
        stats_acc.append(st)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1) / a.steps
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item())
    pairs_total = a.n * (a.n - 1) / 2
    value = pairs_total / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launches, CUDA events inside the library) ----
    def mean(key):
        return sum(s[key] for s in stats_acc) / len(stats_acc)

    st0 = stats_acc[-1]
    peak, peak_src = measured_peak()
    union_kernel = {"bytes": "k_pair_hist", "split": "k_pair_hist_split",
                    "subsets": "k_pair_hist_planes<EpiSubsets>"}.get(os.environ.get("SELB200_UNION", ""), "k_pair_hist_planes")
    k_union = {"name": union_kernel, "ms": mean("ms_union"), "bytes": ALG_BYTES["union"] * st0["pairs_aux"],
               "unit_def": "32768 B (2 x 2^14 one-byte registers, SURVEY.md 8d) per aux-passing pair x pairs_aux"}
    per_pair_filter = {"smh_a": ALG_BYTES["smh_a"], "cb": ALG_BYTES["cb"],
                       "hll_a": 2 * a.aux_bytes, "hll_an": 2 * a.aux_bytes}[a.criterion]
    k_filter = {"name": {"smh_a": "k_tile_filter_smh", "cb": "k_tile_enum", "hll_a": "k_tile_filter_hll",
                         "hll_an": "k_tile_filter_hll"}[a.criterion],
                "ms": mean("ms_filter"), "bytes": per_pair_filter * st0["pairs_cb_shard"],
                "unit_def": f"{per_pair_filter} B per CB-passing pair x pairs_cb(shard)"}
    dom = k_union if k_union["ms"] >= k_filter["ms"] else k_filter
    ach = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
    # the same launch against the integer pipe it is actually bound by.  Essential LOP3 per pair and lane at p=14
    # (8 steps of 64 registers per lane): split kernel 8 x (2 flags + 16 max + 4 group masks + 16 decode + 2 groups x 32)
    # = 816; bit-plane kernel 8 x (20 max + 16 decode + 4 groups x 34) = 1376; its subset form 8 x (20 max + 8 selectors
    # + 8 groups of four x 16) = 1248.  These are the counts with every value group of the window active; the group masks
    # of the C4 workload make the kernels execute about 1200 and 905 (tests/emul/union_lop3_model.py), so `frac` is an
    # upper bound of the essential-LOP3 share.  The LOP3 rate of the SM was measured with tools/ubench/int_pipes.cu
    # (profiles/r01_int_pipes_ubench.txt)
    int_alu = None
    lop3_per_pair = {"k_pair_hist_split": 816.0, "k_pair_hist_planes": 1376.0,
                     "k_pair_hist_planes<EpiSubsets>": 1248.0}.get(dom["name"])
    if lop3_per_pair and dom["ms"] > 0:
        sm_clk = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        lane_ops = lop3_per_pair * 32.0 * st0["pairs_aux"]
        a_int = lane_ops / (dom["ms"] * 1e-3) / sm_clk / n_sm
        int_alu = {"bound": "int_alu", "achieved": a_int, "peak": 63.0, "unit": "LOP3 lane-ops/clk/SM",
                   "frac": a_int / 63.0,
                   "def": f"{lop3_per_pair:.0f} LOP3 per pair and lane (all value groups of the window active: upper bound) "
                          "x 32 lanes x pairs_aux / launch time / SM clock / SMs; peak measured"}
    roofline = {"bound": "hbm", "kernel": dom["name"], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": ncu_traffic(dom["name"]), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom["bytes"], "launch_ms": dom["ms"], "bytes_def": dom["unit_def"],
                "int_alu": int_alu,
                "note": ("frac > 1 is possible: the union kernel reads 6-bit planes (24 KiB per pair, 20 KiB when all values "
                         "are below 32) and L2 serves repeated rows; its binding limit is the integer ALU pipe "
                         "(profiles/r01_ncu_summary.md)"),
                "kernels_ms": {"bounds": mean("ms_bounds"), "filter": mean("ms_filter"), "verify": mean("ms_verify"),
                               "union": mean("ms_union"), "estimate": mean("ms_estimate"), "sort": mean("ms_sort"),
                               "run_total": mean("ms_total")}}

    # ---- e2e: host buffers through the public API ---------------------------------------------------
    e2e = None
    if not a.no_e2e:
        # every rank holds its slice of the file list in pinned host memory (world == 1: everything)
        g0, rows, _per = sdist.slice_rows(a.n, rank, world)
        regs_h = torch.empty((rows, regs_d.shape[1]), dtype=regs_d.dtype, pin_memory=True)
        regs_h.copy_(regs_d[g0:g0 + rows])
        aux_h = None
        if aux_d is not None:
            aux_h = torch.empty((rows, aux_d.shape[1]), dtype=aux_d.dtype, pin_memory=True)
            aux_h.copy_(aux_d[g0:g0 + rows])
        torch.cuda.synchronize()
        sel2 = S.Selection(local, stream=stream)
        sh = None
        if world > 1:
            sdist.setup_gather(sel2)
            sh = sdist.ShardedSketches(a.n, regs_d.shape[1], aux_d.shape[1] if aux_d is not None else 0,
                                       aux_d.dtype if aux_d is not None else None, dev, rank, world)

        phase = {"h2d_gather_load": 0.0, "run_gather_fetch": 0.0}

        def step_e2e():
            t_a = time.perf_counter()
            if world == 1:
                sel2.load(regs_h, aux_h, aux_kind)                       # H2D inside
            else:
                # H2D of the slice piece by piece; each piece is all-gathered over NVLink and digested
                # (validation, histograms, cardinalities, bit planes) while the next one is on the PCIe bus
                sel2.load_device_begin(sh.regs, sh.aux, aux_kind)
                sh.assemble(regs_h, aux_h, on_piece=sel2.load_device_rows)
                sel2.load_end()
            t_b = time.perf_counter()              # load returns after its last device sync
            sel2.run(tau=tau32, criterion=a.criterion, shard=rank, n_shards=world, fetch=False, gather=(world > 1),
                     host_results=(rank == 0))
            out = sel2.result_host()[0].size if rank == 0 else 0
            phase["h2d_gather_load"] += t_b - t_a
            phase["run_gather_fetch"] += time.perf_counter() - t_b
            return out

        e2e_steps = max(1, min(a.steps, 5))
        step_e2e()
        barrier()
        for k_ in phase:
            phase[k_] = 0.0
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(e2e_steps):
            n_out = step_e2e()
        g1.record()
        barrier()
        wall = (time.perf_counter() - t0) / e2e_steps
        dev_ms = g0.elapsed_time(g1) / e2e_steps
        t2 = torch.tensor([max(wall * 1e3, dev_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item())
        h2d = int(regs_d.numel() * regs_d.element_size() + (aux_d.numel() * aux_d.element_size() if aux_d is not None else 0))
        e2e = {"value": pairs_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(out_n * 16 + a.n * 8), "ms_per_step": e2e_ms, "steps": e2e_steps,
               "timer": "max(host wall clock, CUDA events) per step, max over ranks",
               "rank0_phases_ms": {k: v * 1e3 / e2e_steps for k, v in phase.items()}}
        sel2.close()

    cb = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cb = cpu_baseline(a)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u8 registers / u64 buckets, f64 estimator", "data": "synthetic",
                "config": {"workload": workload_name(a), "n": a.n, "pairs": int(pairs_total),
                           "pairs_cb": st0["pairs_cb"], "pairs_aux_rank0": st0["pairs_aux"], "pairs_out": out_n,
                           "bands_x_rows": [st0["n_bands"], st0["n_rows"]], "parallelism": (f"tile-shard x{world}" + (", peer-memory gather to rank 0; e2e: per-rank H2D slices + "
                                                                       "NCCL all-gather" if world > 1 else "")),
                           "l2": "inputs (1.74 GB) larger than L2; no flush needed"},
                "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(sum(s["launches"] for s in stats_acc)),
                "roofline": roofline, "cpu_baseline": cb}
        print(json.dumps(line), file=real_stdout, flush=True)
    sel.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
