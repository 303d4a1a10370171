#!/usr/bin/env python
"""bench.py — sketch-pair comparisons/sec of the all-pairs selection path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's OpenMP selection.cpp

Workload (config C4 of SURVEY.md §8d, the configuration the metric is quoted on): n = 100 000
synthetic bacterial-size genomes ("synth-v1", seed 1002), primary HLL p=14, CB + smh_a with a
1 KiB SuperMinHash auxiliary sketch (m=128), tau = 0.9.  A *step* is one pass of the hot path
over that batch: CB band -> smh_a filter -> HLL-14 union -> Jaccard >= tau -> pair list, sorted,
gathered on rank 0.  metric = [n(n-1)/2] / step time.

  value   : inputs already resident in HBM (sketches generated on the device; with N>1 ranks they
            are broadcast once over NCCL before the timed region), device time by CUDA events on
            the stream the kernels run on, max over ranks.
  e2e     : the same through the public API with HOST buffers: pinned host -> device copy of the
            sketch matrices, per-genome cardinalities + sort, the run, and the device -> host read
            of the pair list, every step (N>1: every rank uploads its slice, NCCL all-gather).
  parity  : outside the timed region.  N=1: the pair list of the benchmark configuration is diffed
            against the stdout of the UNMODIFIED reference binary (oracle/_ref/selection) run on the
            same inputs written as the reference's own gz files, and against the oracle port
            (full-precision Jaccards, stage counts, near-tau list).  N>1: the merged multi-GPU list
            against an unsharded run on rank 0 (whose SHA-256 is the one the N=1 run verified).
  configs : C2, C3 (tau sweep) and C5 of BASELINE.json measured in the same run (resident inputs),
            each with stage counts, per-stage device times, its dominant kernel and (N=1) its own
            oracle diff.
  roofline / kernels / cpu_baseline: see DESIGN.md §7.

One JSON line on stdout (rank 0).  Under torchrun it reads RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sketch_pair_comparisons_per_sec"
UNIT = "pairs/s"
ALG_BYTES = {"cb": 16, "smh_a": 2048, "union": 32768, "emit": 16}   # SURVEY.md §8d, per unit of work
INT_PEAK = 64.0   # alu-pipe lane-ops per clock per SM: 4 SMSPs x 16 lanes (B300_MICROARCH.md "alu-pipe rt_SMSP=2");
#                   tools/ubench/int_pipes.cu measured 63 for LOP3 / PRMT / VIADDMNMX on this pool's B200s
REL_TOL = 1e-6    # north_star: Jaccard within 1e-6 relative; pairs within 1e-6 of tau listed separately


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
    ap.add_argument("--ref-budget-s", type=float, default=900.0,
                    help="--impl reference: wall-clock budget for the timed full-configuration passes")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reference pass (and with it the parity block)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2 / C3 / C5 block")
    ap.add_argument("--no-config-parity", action="store_true", help="configs block without its oracle diffs")
    ap.add_argument("--cfg-steps", type=int, default=5)
    return ap.parse_args()


class Cfg:
    """One workload: what both arms put into `config` (identical dicts) and what the generators need."""

    def __init__(self, n, seed, criterion, tau, aux_bytes):
        self.n, self.seed, self.criterion, self.tau = int(n), int(seed), criterion, float(tau)
        self.aux_bytes = int(aux_bytes) if criterion != "cb" else 0

    @property
    def pairs(self):
        return self.n * (self.n - 1) // 2

    def name(self):
        return (f"synth-v1 seed {self.seed}: n={self.n} genomes, HLL p=14, CB + {self.criterion}"
                + (f" (aux {self.aux_bytes} B)" if self.criterion != "cb" else "") + f", tau={self.tau}")

    def config(self):
        return {"workload": self.name(), "n": self.n, "pairs": self.pairs, "criterion": self.criterion, "tau": self.tau,
                "aux_bytes": self.aux_bytes, "seed": self.seed, "p": 14,
                "l2": "inputs larger than L2 at n >= 10k (n x 16 KiB registers); no flush needed"}


def other_configs():
    """BASELINE.json configs[1], [2], [4]; C4 is the headline workload."""
    out = [("C2 cb", Cfg(10_000, 1001, "cb", 0.9, 0))]
    out += [(f"C3 smh_a tau={t:.2f}", Cfg(10_000, 1001, "smh_a", t, 1024)) for t in (0.70, 0.75, 0.80, 0.85, 0.90, 0.95)]
    out += [("C5 hll_a p_aux=10", Cfg(50_000, 1003, "hll_a", 0.9, 1024)), ("C5 hll_an p_aux=10", Cfg(50_000, 1003, "hll_an", 0.9, 1024)),
            ("C5 hll_a p_aux=8", Cfg(50_000, 1003, "hll_a", 0.9, 256)), ("C5 hll_an p_aux=8", Cfg(50_000, 1003, "hll_an", 0.9, 256))]
    return out


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
# the reference's own OpenMP selection.cpp (oracle/_ref/selection) on the workload written as its gz files
# ----------------------------------------------------------------------------------------------
def file_names(n):
    return [f"g{i:06d}.fna.gz" for i in range(n)]


def host_inputs(cfg: Cfg):
    """(regs uint8[n][2^14], aux or None) from the host generator (bit-identical to the device one, tested)."""
    from cuda_selection_criteria_b200 import synth
    plan = synth.make_plan(cfg.n, cfg.seed)
    regs = synth.hll(plan, 14)
    aux = None
    if cfg.criterion == "smh_a":
        aux = synth.smh(plan, cfg.aux_bytes // 8)
    elif cfg.criterion in ("hll_a", "hll_an"):
        aux = synth.hll(plan, cfg.aux_bytes.bit_length() - 1, synth.TAG_AUX_HLL)
    return regs, aux


_W = {}


def _write_range(rg):
    from cuda_selection_criteria_b200 import sketch_io
    a, b = rg
    td, names, regs, smh, aux, pa = _W["td"], _W["names"], _W["regs"], _W["smh"], _W["aux"], _W["pa"]
    for i in range(a, b):
        base = os.path.join(td, names[i])
        sketch_io.write_hll(base + ".hll", regs[i], 14, level=1)
        if smh is not None:
            sketch_io.write_smh(base + ".smh" + str(smh.shape[1]), smh[i], level=1)
        if aux is not None:
            sketch_io.write_hll(base + ".hll_" + str(pa), aux[i], pa, level=1)
    return b - a


def write_files(cfg: Cfg, regs, aux, td: str):
    """The reference's on-disk layout (sketch_io) for the whole workload, written by forked workers (the parent
    must not have touched CUDA yet: the reference leg runs before the device is initialised)."""
    import multiprocessing as mp
    n = cfg.n
    smh = auxh = None
    pa = 0
    if cfg.criterion == "smh_a":
        smh = aux
    elif cfg.criterion in ("hll_a", "hll_an"):
        auxh, pa = aux, cfg.aux_bytes.bit_length() - 1
    else:
        smh = np.full((n, 1), 42, np.uint64)     # CB only: constant .smh1 -> smh_a always passes
    _W.update(td=td, names=file_names(n), regs=regs, smh=smh, aux=auxh, pa=pa)
    nw = os.cpu_count() or 1
    step = max(1, (n + nw * 8 - 1) // (nw * 8))
    with mp.get_context("fork").Pool(nw) as pool:
        done = sum(pool.map(_write_range, [(a, min(n, a + step)) for a in range(0, n, step)]))
    assert done == n
    with open(os.path.join(td, "list.txt"), "w") as f:
        f.write("\n".join(_W["names"]) + "\n")
    _W.clear()


def ref_cmd(cfg: Cfg, cores: int, tau):
    ref = os.path.join(ROOT, "oracle", "_ref", "selection")
    crit = "smh_a" if cfg.criterion == "cb" else cfg.criterion
    aux_bytes = 8 if cfg.criterion == "cb" else cfg.aux_bytes
    return [ref, "-l", "list.txt", "-t", str(cores), "-a", str(aux_bytes), "-c", crit, "-h", str(tau)]


def ref_pass(cfg: Cfg, td: str, cores: int, tau=None):
    """(wall seconds, stdout bytes) of one run of the unmodified binary over the whole workload."""
    t0 = time.perf_counter()
    out = subprocess.run(ref_cmd(cfg, cores, cfg.tau if tau is None else tau), cwd=td, capture_output=True, check=True).stdout
    return time.perf_counter() - t0, out


class ReferenceLeg:
    """Materialises the workload once and runs the reference on ALL of it.  compare phase = wall(-h tau) minus
    wall(-h 2.0) (CB breaks every row at once: load + sort only), BASELINE.md §4.2."""

    def __init__(self, cfg: Cfg):
        self.cfg = cfg
        self.cores = os.cpu_count() or 1
        self.have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "selection"))
        self.td = tempfile.mkdtemp(prefix="selb200_ref_")
        self.lines = None
        self.t_prep = self.t_load = 0.0
        self.regs = self.aux = None

    def prepare(self):
        t0 = time.perf_counter()
        self.regs, self.aux = host_inputs(self.cfg)
        if self.have_ref:
            write_files(self.cfg, self.regs, self.aux, self.td)
            self.t_prep = time.perf_counter() - t0
            self.t_load = min(ref_pass(self.cfg, self.td, self.cores, tau=2.0)[0] for _ in range(2))
        else:
            self.t_prep = time.perf_counter() - t0

    def one_pass(self):
        """compare-phase seconds of one full pass; keeps the reference's stdout lines of the first one."""
        if self.have_ref:
            wall, out = ref_pass(self.cfg, self.td, self.cores)
            if self.lines is None:
                self.lines = out.decode().splitlines()
            return max(wall - self.t_load, 1e-9), wall
        # no reference binary on this machine (it is built where /root/reference exists): the oracle port
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_api as O
        t0 = time.perf_counter()
        res = O.select(self.regs, 14, self.cfg.criterion, np.float32(self.cfg.tau),
                       aux=self.aux if self.cfg.criterion != "cb" else None, threads=self.cores)
        wall = time.perf_counter() - t0
        if self.lines is None:
            self.lines = O.format_lines(file_names(self.cfg.n), res)
        return wall, wall

    def cli_pass(self, runs=2):
        """The process-level drop-in: this repo's own `selection` binary on the SAME gz files and flags
        as the reference binary, whole process timed (CUDA start-up, gunzip, upload, compare, print), stdout held
        against the reference's byte for byte.  Outside every timed region of the metric; reported as `cli`."""
        exe = os.path.join(ROOT, "cuda_selection_criteria_b200", "bin", "selection")
        if not (self.have_ref and os.path.exists(exe) and self.lines is not None):
            return None
        cmd = [exe] + ref_cmd(self.cfg, self.cores, self.cfg.tau)[1:]
        walls, out, err, phases = [], b"", None, []
        for _ in range(runs):
            t0 = time.perf_counter()
            r = subprocess.run(cmd + ["-g"], cwd=self.td, capture_output=True)     # -g: phase times on stderr, stdout unchanged
            walls.append(round(time.perf_counter() - t0, 3))
            if r.returncode != 0:
                err = f"exit {r.returncode}: " + r.stderr.decode(errors="replace")[-200:]
                break
            out = r.stdout
            phases += [ln[len("selb200: "):] for ln in r.stderr.decode(errors="replace").splitlines() if ln.startswith("selb200: host ms")]
        got = out.decode().splitlines()
        return {"cmd": "cuda_selection_criteria_b200/bin/selection " + " ".join(cmd[1:]), "wall_s": walls,
                "reference_cmd": "oracle/_ref/selection " + " ".join(cmd[1:]), "lines": len(got),
                "stdout_identical_to_reference": err is None and got == self.lines, "stdout_sha256_16": lines_sha(got)[:16],
                "phases": phases,
                "error": err}

    def drop_files(self):
        shutil.rmtree(self.td, ignore_errors=True)

    def baseline(self, secs, walls):
        t = sum(secs) / len(secs)
        kind = "reference" if self.have_ref else "port"
        return {"value": self.cfg.pairs / t, "unit": UNIT, "cores": self.cores, "kind": kind,
                "sample": (f"the whole workload ({self.cfg.pairs} pairs, all {self.cfg.n} genomes), {len(secs)} full pass(es) of "
                           + (f"oracle/_ref/selection -t {self.cores} on the reference's own gz files; compare phase = "
                              f"wall(-h {self.cfg.tau}) - wall(-h 2.0)" if self.have_ref else
                              f"the oracle port (oracle/liboracle.so, {self.cores} threads, incl. cardinalities)")),
                "seconds_per_pass": t, "passes": len(secs), "wall_tau_s": [round(w, 2) for w in walls],
                "wall_load_only_s": round(self.t_load, 2), "prep_s": round(self.t_prep, 1),
                "lines": len(self.lines) if self.lines is not None else None}


def run_reference_arm(a, cfg: Cfg, rank: int, out):
    """--impl reference: the unmodified binary on the SAME configuration (all n genomes), every step one full pass.
    A full pass takes minutes of host-core time at n = 100k (std::map<string> lookups and by-value vector copies per
    CB pair, src/selection.cpp:279-284), so when warmup + steps passes do not fit --ref-budget-s the arm times as many
    full passes as fit (at least one after one warm-up pass) and says so in `steps` / `steps_requested`."""
    if rank != 0:
        return
    leg = ReferenceLeg(cfg)
    try:
        leg.prepare()
        t0 = time.perf_counter()
        first_cmp, first_wall = leg.one_pass()                 # warm-up pass 1 (page cache, and the pair list)
        per_pass = time.perf_counter() - t0
        fit = int(max(0.0, a.ref_budget_s - per_pass) // max(per_pass, 1e-3))
        if fit >= (a.warmup - 1) + a.steps:
            n_warm, n_steps = a.warmup, a.steps
        else:
            n_warm, n_steps = 1, max(1, min(a.steps, fit))
        for _ in range(n_warm - 1):
            leg.one_pass()
        secs, walls = [], []
        for _ in range(n_steps):
            c, w = leg.one_pass()
            secs.append(c); walls.append(w)
        cb = leg.baseline(secs, walls)
    finally:
        leg.drop_files()
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": n_steps, "warmup": n_warm, "steps_requested": a.steps, "warmup_requested": a.warmup,
            "ms_per_step": cb["seconds_per_pass"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 registers / u64 buckets, f64 estimator", "data": "synthetic", "config": cfg.config(),
            "parallelism": f"host OpenMP, {cb['cores']} threads", "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if (n_warm, n_steps) != (a.warmup, a.steps):
        line["note"] = (f"{a.warmup}+{a.steps} full passes of {per_pass:.0f} s do not fit the {a.ref_budget_s:.0f} s budget: "
                        f"{n_warm} warm-up + {n_steps} timed full-configuration passes were run instead of a smaller sample")
    print(json.dumps(line), file=out, flush=True)


# ----------------------------------------------------------------------------------------------
# parity helpers (outside every timed region)
# ----------------------------------------------------------------------------------------------
def lines_sha(lines):
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()


def diff_reference_lines(got_lines, ref_lines, tau):
    """Pair-set diff of `name name jaccard` lines (run_comparison_experiment.sh:36-52 joins on the two names)."""
    def keyed(ls):
        d = {}
        for ln in ls:
            p = ln.rsplit(" ", 1)
            d[p[0]] = p[1]
        return d
    g, r = keyed(got_lines), keyed(ref_lines)
    common = [k for k in r if k in g]
    max_abs = max((abs(float(g[k]) - float(r[k])) for k in common), default=0.0)
    return {"against": "stdout of the unmodified reference binary (oracle/_ref/selection) on the same inputs, all n genomes",
            "pairs_reference": len(r), "pairs_ours": len(g), "pairs_missing": len(r) - len(common),
            "pairs_extra": len(g) - len(common), "lines_identical_in_order": got_lines == ref_lines,
            "max_abs_diff_printed_jaccard": max_abs,
            "printed_precision": "std::to_string: 6 decimals, so identical strings bound |dJ| by 1e-6 absolute; "
                                 "the full-precision comparison is the `oracle_port` block"}


def diff_oracle(res, ora, tau):
    """Full-precision diff against the oracle port: pair set, stage counts, relative Jaccard error, near-tau list."""
    got = {(int(i), int(k)): float(j) for i, k, j in zip(res.i, res.k, res.jaccard)}
    want = {(int(i), int(k)): float(j) for i, k, j in zip(ora["i"], ora["k"], ora["jaccard"])}
    near_j = {(int(i), int(k)): float(j) for i, k, j in zip(res.near_i, res.near_k, res.near_jaccard)}
    near = sorted(near_j)
    missing = [p for p in want if p not in got and p not in near]
    extra = [p for p in got if p not in want and p not in near]
    common = [p for p in want if p in got]
    rel = max((abs(got[p] - want[p]) / max(abs(want[p]), 1e-300) for p in common), default=0.0)
    st = res.stats
    return {"against": "oracle/liboracle.so (CPU restatement pinned to the reference's goldens), same inputs, all n genomes",
            "pairs_oracle": len(want), "pairs_ours": len(got), "pairs_missing": len(missing), "pairs_extra": len(extra),
            "order_identical": list(zip(res.i.tolist(), res.k.tolist())) == list(zip(ora["i"].tolist(), ora["k"].tolist())),
            "max_rel_jaccard": rel, "rel_tol": REL_TOL, "jaccard_bits_identical": bool(
                len(got) == len(want) and np.array_equal(res.jaccard.view(np.int64), ora["jaccard"].view(np.int64))),
            "pairs_cb_equal": st["pairs_cb"] == ora["stage"][1], "pairs_aux_equal": st["pairs_aux"] == ora["stage"][2],
            "pairs_out_equal": st["pairs_out"] == ora["stage"][3], "cards_truncated_equal": bool(np.array_equal(
                res.cards_sorted.astype(np.uint64), ora["cards_sorted"].astype(np.uint64))),
            "near_tau": [[i, k, near_j[(i, k)], (i, k) in got] for i, k in near][:64], "near_tau_count": len(near),
            "near_tau_fields": "sorted row i, sorted column k, Jaccard, emitted",
            "near_tau_equal_oracle": near == sorted(zip(ora.get("near_i", np.zeros(0, np.int32)).tolist(),
                                                         ora.get("near_k", np.zeros(0, np.int32)).tolist()))}


def parity_ok(p):
    return (p["pairs_missing"] == 0 and p["pairs_extra"] == 0 and p.get("max_rel_jaccard", 0.0) <= REL_TOL
            and p.get("pairs_cb_equal", True) and p.get("pairs_aux_equal", True) and p.get("cards_truncated_equal", True))


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


def kernel_table(cfg: Cfg, st, mean, sm_clk_hz, n_sm, peak_hbm, n_shards=1):
    """One entry per stage of the run (this rank's launches, CUDA events inside the library): time, share of the run,
    the roof that binds it and the achieved fraction.  Integer-pipe entries count ESSENTIAL lane-operations (the
    instructions the algorithm needs, not the ones issued) against the alu-pipe rate of the SM; byte entries count
    SURVEY.md §8d's algorithmic bytes against the measured HBM rate."""
    run_ms = mean("ms_total")
    rows = []

    def add(stage, kernel, ms, bound, work, per_unit, unit_def, note=None):
        if ms <= 0:
            return
        e = {"stage": stage, "kernel": kernel, "ms": ms, "share_of_run": ms / run_ms if run_ms > 0 else None, "bound": bound}
        if bound == "int_alu":
            ach = work * per_unit / (ms * 1e-3) / sm_clk_hz / n_sm
            e.update(achieved=ach, peak=INT_PEAK, unit="essential alu lane-ops/clk/SM", frac=ach / INT_PEAK)
        else:
            ach = work * per_unit / (ms * 1e-3) / 1e9
            e.update(achieved=ach, peak=peak_hbm, unit="GB/s (algorithmic bytes)", frac=ach / peak_hbm)
        e["def"] = unit_def
        if note:
            e["note"] = note
        rows.append(e)

    p_cb, p_cand, p_aux, p_out = st["pairs_cb_shard"], st["pairs_cand"], st["pairs_aux"], st["pairs_out"]
    add("bounds", "k_cb_bounds + k_rowblock_span + scan + k_tile_table", mean("ms_bounds"), "hbm", cfg.n, 16.0 * 20,
        "16 B per tested pair x ~20 binary-search probes per row", "latency-bound: four small launches")
    smh_mode = os.environ.get("SELB200_SMHFILTER", "")
    if cfg.criterion == "smh_a" and (smh_mode == "join" or (smh_mode != "tiles" and n_shards < 4)):
        nb = st["n_bands"]
        # equality join: n x bands elements bucketed by a counting sort and walked once, plus the exact bucket compare of
        # the candidates
        add("filter", "k_smh_sigkeys + scan + k_smh_scatter + k_smh_join_expand + k_smh_join", mean("ms_filter"), "hbm", cfg.n * nb,
            8.0 * st["n_rows"] + 24.0,
            f"{cfg.n} x {nb} (genome, band) elements: the band's {st['n_rows']} buckets read once (8 B each) + key, rank and bucket slot "
            "written and read (24 B); work is O(n x bands + matches), not O(P_cb x bands)",
            "latency-bound: six small launches (keys, two of the scan, scatter, expansion, walk), dependent loads per item")
    elif cfg.criterion == "smh_a":
        nb = st["n_bands"]
        add("filter", "k_smh_signatures + k_tile_filter_smh", mean("ms_filter"), "int_alu", p_cb, float((nb + 1) // 2),
            f"one packed min/add lane-operation (VIADDMNMX.U16x2) per two LSH bands and pair: {(nb + 1) // 2} per CB pair of the shard "
            "(a thread owns an 8x8 block of pairs, so one lane-operation serves one pair)")
        add("verify", "k_smh_verify", mean("ms_verify"), "hbm", p_cand, 2.0 * 8 * st["n_rows"] + 16,
            "2 x 8 x n_rows bucket bytes + 16 B per candidate", "latency-bound: thread per candidate, dependent loads")
    elif cfg.criterion in ("hll_a", "hll_an"):
        regs_aux = cfg.aux_bytes
        mode = os.environ.get("SELB200_HLLFILTER", "")
        if mode in ("onepass", "bytes") or regs_aux < 64:
            add("filter", "k_tile_filter_hll_planes" if mode == "onepass" else "k_tile_filter_hll", mean("ms_filter"), "int_alu", p_cb,
                hll_essential_lop3(regs_aux),
                f"essential LOP3 per CB pair (thread per pair, so lane-op = op): {hll_essential_lop3(regs_aux):.0f} for "
                f"{regs_aux} auxiliary registers (max 12 + subset counting of 5 four-value groups: 6 selectors + 80, per 64 registers)")
        else:
            # pass A stops a warp step once its 32 pairs are decided, so the work it needs is what it read: the library
            # counts the executed steps (stats.filter_steps: 64 registers x 32 pairs each)
            steps = st["filter_steps"]
            full = p_cb / 32.0 * (regs_aux / 64.0)
            add("filter", "k_tile_filter_hll_bound", mean("ms_filter"), "int_alu", steps * 32.0, HLL_BOUND_LOP3,
                f"essential LOP3 per executed step and pair: {HLL_BOUND_LOP3:.0f} per 64 registers (max 2x12 + 4 selectors + "
                f"3 counted four-value groups x 16); {steps} warp steps executed = {steps / max(full, 1):.3f} of the "
                f"{full:.0f} a full read of every CB pair's sketches would take (bound decided after 1/2 or 3/4 of the registers)")
            add("verify", "k_hll_verify", mean("ms_verify"), "int_alu", p_cand, hll_essential_lop3(regs_aux),
                f"essential LOP3 per candidate: {hll_essential_lop3(regs_aux):.0f} (full sketch) ; then the fp64 Ertl MLE per candidate",
                "fp64-latency-bound: thread per candidate, secant loop on a shared-memory histogram column")
    else:
        add("filter", "k_tile_enum", mean("ms_filter"), "hbm", p_cb, 16.0, "16 B per CB pair (emit the band)")
    lop3 = union_lop3_per_pair()
    add("union", union_kernel_name(), mean("ms_union"), "int_alu", p_aux, lop3 * 32.0,
        f"{lop3:.0f} essential LOP3 per pair and lane (all value groups of the 32-value window active: upper bound) x 32 lanes")
    if rows and rows[-1]["stage"] == "union" and union_kernel_name().endswith("<EpiSubsets>"):
        # The subset kernel counts one mask of four with POPC (XU pipe, 16 lanes / clk / SM) and three with carry-save LOP3
        # (ALU pipe, 64): both pipes carry about the same time, so both are stated — with every group of the window active
        # (upper bound, the figure above) and with the groups the C4 workload executes (per-eighth limits, value range of
        # the pairs: tests/emul/union_lop3_model.py on the same generator).  frac = the busier pipe on the EXECUTED counts.
        e = rows[-1]
        clk_per_pair = mean("ms_union") * 1e-3 * sm_clk_hz * n_sm / max(p_aux, 1)
        pipes = {}
        for name, rate, all_groups, executed in (("alu", INT_PEAK, lop3, UNION_EXECUTED["lop3"]), ("xu", XU_PEAK, UNION_ALL["popc"], UNION_EXECUTED["popc"])):
            pipes[name] = {"rate_lane_ops_per_clk_sm": rate,
                           "all_groups": {"ops_per_pair_lane": all_groups, "frac": all_groups * 32.0 / rate / clk_per_pair},
                           "executed_c4": {"ops_per_pair_lane": executed, "frac": executed * 32.0 / rate / clk_per_pair}}
        busier = max(pipes, key=lambda k: pipes[k]["executed_c4"]["frac"])
        e["pipes"] = pipes
        e["clk_per_pair_per_sm"] = clk_per_pair
        if cfg.n == 100_000 and cfg.criterion == "smh_a":
            e["frac_all_groups_alu"] = e["frac"]
            e["frac"] = pipes[busier]["executed_c4"]["frac"]
            e["achieved"] = e["frac"] * pipes[busier]["rate_lane_ops_per_clk_sm"]
            e["peak"] = pipes[busier]["rate_lane_ops_per_clk_sm"]
            e["unit"] = f"executed {busier}-pipe lane-ops/clk/SM"
            e["def"] = (f"{pipes[busier]['executed_c4']['ops_per_pair_lane']:.0f} {'POPC' if busier == 'xu' else 'LOP3'} per pair and lane that the "
                        f"C4 workload executes (tests/emul/union_lop3_model.py) x 32 lanes on the {busier} pipe "
                        f"({pipes[busier]['rate_lane_ops_per_clk_sm']:.0f} lanes/clk/SM); `pipes` has both pipes, also with all groups active")
    add("estimate", "k_estimate_emit", mean("ms_estimate"), "hbm", p_aux, 256.0 + 16,
        "256 B histogram row + 16 B descriptor per aux-passing pair", "fp64-latency-bound: thread per histogram, Ertl MLE secant loop")
    add("sort", "k_rowsort_* + D2H", mean("ms_sort"), "hbm", max(p_out, 1), 16.0 * 4, "16 B per emitted pair, four passes",
        "latency-bound: four small launches + the copy to pinned host memory")
    return rows


def union_kernel_name():
    return {"bytes": "k_pair_hist", "planes": "k_pair_hist_planes<one-hot>"}.get(os.environ.get("SELB200_UNION", ""),
                                                                                "k_pair_hist_planes<EpiSubsets>")


def union_lop3_per_pair():
    # essential LOP3 per pair and lane at p=14 (8 steps of 64 registers per lane): one-hot form 8 x (20 max + 16 decode +
    # 4 groups x 34) = 1376; subset form 8 x (20 max + 8 selectors + 8 groups of four x 14) = 1120 (one mask of four is
    # counted with POPC, kernels/union_planes.inl PL_DIRECT).  With the group masks of the C4 workload the kernels execute
    # about 1200 and 738 (tests/emul/union_lop3_model.py).
    return {"planes": 1376.0, "bytes": 0.0}.get(os.environ.get("SELB200_UNION", ""), 1120.0)


XU_PEAK = 16.0          # POPC lanes per clock per SM (tools/ubench/int_pipes.cu: 16)
UNION_ALL = {"popc": 8 * 8 * 5 + 24.0}                 # all eight groups: 5 POPC per group and step, 24 in the epilogue
UNION_EXECUTED = {"lop3": 738.0, "popc": 189.0 + 24.0}   # C4 pairs, tests/emul/union_lop3_model.py (4.73 groups per step)


HLL_BOUND_LOP3 = 24.0 + 4 + 3 * 16      # pass A of the two-pass hll filter, per step of 64 auxiliary registers


def hll_essential_lop3(regs_aux):
    return (regs_aux / 64.0) * (24 + 6 + 5 * 16)


def main():
    a = parse()
    real_stdout = _protect_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = Cfg(a.n, a.seed, a.criterion, a.tau, a.aux_bytes)
    if a.impl == "reference":
        run_reference_arm(a, cfg, rank, real_stdout)
        return

    # ---- reference leg first (rank 0 at N=1 only), before this process touches CUDA: the file writers fork -------
    cb = None
    ref_lines = None
    cli = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        leg = ReferenceLeg(cfg)
        try:
            leg.prepare()
            c, w = leg.one_pass()
            cb = leg.baseline([c], [w])
            ref_lines = leg.lines
            cli = leg.cli_pass()
            if cli is not None:
                cli["reference_wall_s"] = round(w, 3)
        finally:
            leg.drop_files()
        del leg

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

    def aux_kind_of(c: Cfg):
        return {"cb": AUX_NONE, "smh_a": AUX_SMH, "hll_a": AUX_HLL, "hll_an": AUX_HLL}[c.criterion]

    def resident_inputs(c: Cfg):
        """sketches generated on rank 0's device, broadcast over NCCL"""
        plan = synth.make_plan(c.n, c.seed)
        kind = aux_kind_of(c)
        if rank == 0:
            regs_d = synth.hll(plan, 14, device=local)
            aux_d = None
            if kind == AUX_SMH:
                aux_d = synth.smh(plan, c.aux_bytes // 8, device=local)
            elif kind == AUX_HLL:
                aux_d = synth.hll(plan, c.aux_bytes.bit_length() - 1, synth.TAG_AUX_HLL, device=local)
        else:
            regs_d = torch.empty((c.n, 1 << 14), dtype=torch.uint8, device=dev)
            aux_d = None
            if kind == AUX_SMH:
                aux_d = torch.empty((c.n, c.aux_bytes // 8), dtype=torch.int64, device=dev)
            elif kind == AUX_HLL:
                aux_d = torch.empty((c.n, c.aux_bytes), dtype=torch.uint8, device=dev)
        if world > 1:
            sdist.broadcast_sketches(regs_d, aux_d, src=0)
        torch.cuda.synchronize()
        return regs_d, aux_d, kind

    regs_d, aux_d, aux_kind = resident_inputs(cfg)

    # one explicit stream for everything: torch copies and NCCL collectives issued under it and the
    # library's kernels (which run on the stream handed to the context) are then ordered on the device
    # (the legacy default stream has handle 0, which the C-ABI reads as "use a stream of your own")
    main_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(main_stream)
    stream = main_stream.cuda_stream
    assert stream != 0
    sel = S.Selection(local, stream=stream)
    sel.load(regs_d, aux_d, aux_kind)
    if world > 1:
        sdist.setup_gather(sel)        # rank 0's landing zone, mapped into every rank (CUDA IPC)

    def timed_steps(sel_, c: Cfg, warmup, steps, sample_clocks):
        """W warm-up + K timed steps of the resident path; (ms per step max over ranks, stats of every step, clocks)"""
        tau32 = np.float32(c.tau)
        acc = []

        def step():
            # with several ranks every rank's kernels store its pairs into rank 0's memory over NVLink
            # (peer-memory gather inside selb200_run); rank 0 sorts the merged list and its run ends with the
            # D2H of that list into pinned host memory — inside the metric
            res = sel_.run(tau=tau32, criterion=c.criterion, shard=rank, n_shards=world, fetch=False,
                           gather=(world > 1), host_results=(rank == 0))
            return res.stats, (sel_.result_host()[0].size if rank == 0 else 0)

        for _ in range(warmup):
            step()
        barrier()
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        out_n = 0
        for _ in range(steps):
            st, out_n = step()
            acc.append(st)
        ev1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        t_ms = torch.tensor([ev0.elapsed_time(ev1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return float(t_ms.item()), acc, out_n, clocks

    ms_step, stats_acc, out_n, clocks = timed_steps(sel, cfg, a.warmup, a.steps, True)
    value = cfg.pairs / (ms_step * 1e-3)

    def mean_of(acc):
        return lambda key: sum(s[key] for s in acc) / len(acc)

    st0 = stats_acc[-1]
    peak, peak_src = measured_peak()
    sm_clk = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    kernels = kernel_table(cfg, st0, mean_of(stats_acc), sm_clk, n_sm, peak, world)
    dom = max(kernels, key=lambda e: e["ms"])
    dom_alg_bytes = {"union": ALG_BYTES["union"] * st0["pairs_aux"],
                     "filter": {"smh_a": ALG_BYTES["smh_a"], "cb": ALG_BYTES["cb"], "hll_a": 2 * cfg.aux_bytes,
                                "hll_an": 2 * cfg.aux_bytes}[cfg.criterion] * st0["pairs_cb_shard"]}.get(dom["stage"], 0)
    hbm_ach = dom_alg_bytes / (dom["ms"] * 1e-3) / 1e9
    # the contract's block for the DOMINANT kernel.  Its binding roof is the integer ALU pipe (ncu: profiles/), so
    # bound / achieved / peak / frac are stated against that pipe; the algorithmic-byte figure of SURVEY.md §8d against
    # the measured HBM rate sits beside it under "hbm" (it can exceed 1: bit planes move 20-24 KiB per pair, not 32,
    # and L2 serves repeated rows — see `traffic`)
    roofline = {"bound": dom["bound"], "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": dom["peak"],
                "unit": dom["unit"], "frac": dom["frac"], "traffic": ncu_traffic(dom["kernel"]), "launch_ms": dom["ms"],
                "def": dom["def"],
                "peak_source": (("xu pipe (POPC): 4 SMSPs x 4 lanes per clock (tools/ubench/int_pipes.cu measured 16 on this pool)"
                                 if "xu-pipe" in dom["unit"] else
                                 "alu pipe: 4 SMSPs x 16 lanes per clock (B300_MICROARCH.md 'alu-pipe rt_SMSP=2'; "
                                 "tools/ubench/int_pipes.cu measured 63 on this pool)") if dom["bound"] == "int_alu" else peak_src),
                "pipes": dom.get("pipes"), "frac_all_groups_alu": dom.get("frac_all_groups_alu"),
                "hbm": {"achieved": hbm_ach, "peak": peak, "unit": "GB/s", "frac": hbm_ach / peak, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": dom_alg_bytes,
                        "note": "SURVEY.md 8d algorithmic bytes / launch time; not the binding roof when frac > 1"},
                "kernels_ms": {k: mean_of(stats_acc)(f"ms_{k}") for k in ("bounds", "filter", "verify", "union", "estimate", "sort")}
                | {"run_total": mean_of(stats_acc)("ms_total")}}

    # ---- parity of the benchmark configuration (untimed) ---------------------------------------------
    parity = None
    names = file_names(cfg.n)
    full = sel.run(tau=np.float32(cfg.tau), criterion=cfg.criterion, shard=rank, n_shards=world, gather=(world > 1),
                   fetch=(rank == 0))
    if rank == 0:
        got_lines = S.format_lines(names, full)
        parity = {"lines_sha256": lines_sha(got_lines), "pairs": len(got_lines)}
    if world > 1:
        # the merged multi-GPU list against an unsharded run of the same context on rank 0
        if rank == 0:
            solo = sel.run(tau=np.float32(cfg.tau), criterion=cfg.criterion)
            same = (np.array_equal(solo.i, full.i) and np.array_equal(solo.k, full.k)
                    and np.array_equal(solo.jaccard.view(np.int64), full.jaccard.view(np.int64)))
            parity["multi_gpu"] = {"against": "unsharded run on rank 0 (the list whose SHA-256 the N=1 run checks against the reference)",
                                   "pairs_single": int(solo.i.size), "pairs_merged": int(full.i.size),
                                   "identical_pairs_order_and_jaccard_bits": bool(same),
                                   "single_gpu_lines_sha256": lines_sha(S.format_lines(names, solo))}
            parity["ok"] = bool(same)
        barrier()
    elif rank == 0 and ref_lines is not None:
        parity["reference_binary" if cb["kind"] == "reference" else "oracle_lines"] = diff_reference_lines(got_lines, ref_lines, cfg.tau)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_api as O
        regs_h = regs_d.cpu().numpy()
        aux_h = None if aux_d is None else (aux_d.cpu().numpy().view(np.uint64) if aux_kind == AUX_SMH else aux_d.cpu().numpy())
        ora = O.select(regs_h, 14, cfg.criterion, np.float32(cfg.tau), aux=aux_h, threads=os.cpu_count() or 1)
        parity["oracle_port"] = diff_oracle(full, ora, cfg.tau)
        del regs_h, aux_h, ora
        parity["ok"] = all(parity_ok(parity[k]) for k in ("reference_binary", "oracle_lines", "oracle_port") if k in parity)
    del full

    # ---- e2e: host buffers through the public API ---------------------------------------------------
    e2e = None
    if not a.no_e2e:
        # every rank holds its slice of the file list in pinned host memory (world == 1: everything)
        g0, rows, _per = sdist.slice_rows(cfg.n, rank, world)
        regs_h = torch.empty((rows, regs_d.shape[1]), dtype=regs_d.dtype, pin_memory=True)
        regs_h.copy_(regs_d[g0:g0 + rows])
        aux_h = None
        if aux_d is not None:
            aux_h = torch.empty((rows, aux_d.shape[1]), dtype=aux_d.dtype, pin_memory=True)
            aux_h.copy_(aux_d[g0:g0 + rows])
        torch.cuda.synchronize()
        # what the link itself delivers: the same pinned buffer copied alone (best of 3), every rank at once
        link = []
        scratch = torch.empty_like(regs_h, device=dev)
        for _ in range(3):
            barrier()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            scratch.copy_(regs_h, non_blocking=True)
            l1.record()
            torch.cuda.synchronize()
            link.append(regs_h.numel() / (l0.elapsed_time(l1) * 1e-3) / 1e9)
        del scratch
        link_gbs = torch.tensor([max(link)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(link_gbs, op=dist.ReduceOp.SUM)
        sel2 = S.Selection(local, stream=stream)
        sh = None
        if world > 1:
            sdist.setup_gather(sel2)
            sh = sdist.ShardedSketches(cfg.n, regs_d.shape[1], aux_d.shape[1] if aux_d is not None else 0,
                                       aux_d.dtype if aux_d is not None else None, dev, rank, world)

        phase = {"h2d_gather_load": 0.0, "run_gather_fetch": 0.0}
        tau32 = np.float32(cfg.tau)

        def step_e2e():
            t_a = time.perf_counter()
            if world == 1:
                sel2.load(regs_h, aux_h, aux_kind)                       # H2D inside
            else:
                # H2D of the slice piece by piece; each piece is all-gathered over NVLink and digested
                # (validation, histograms, cardinalities, bit planes) while the next one is on the PCIe bus
                sel2.load_device_begin(sh.regs, sh.aux, aux_kind)
                sh.assemble(regs_h, aux_h, on_piece=sel2.load_device_rows, on_piece_packed=sel2.load_device_rows_packed)
                sel2.load_end()
            t_b = time.perf_counter()              # load returns after its last device sync
            sel2.run(tau=tau32, criterion=cfg.criterion, shard=rank, n_shards=world, fetch=False, gather=(world > 1),
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
        g0e, g1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0e.record()
        for _ in range(e2e_steps):
            step_e2e()
        g1e.record()
        barrier()
        wall = (time.perf_counter() - t0) / e2e_steps
        dev_ms = g0e.elapsed_time(g1e) / e2e_steps
        t2 = torch.tensor([max(wall * 1e3, dev_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item())
        host_bytes = int(regs_d.numel() * regs_d.element_size() + (aux_d.numel() * aux_d.element_size() if aux_d is not None else 0))
        h2d = host_bytes
        h2d_mode = os.environ.get("SELB200_H2D", "")
        packed_h2d = (h2d_mode == "packed") or (h2d_mode != "raw" and world == 1)
        load_info = None
        if world == 1:
            # what crossed the bus in the last step, as the library counted it: register pieces packed (csrc/hostpack.h) or
            # raw (whenever the link had run out of work), + the auxiliary sketches as they are
            load_info = sel2.load_info()
            h2d = int(load_info["h2d_register_bytes"]) + (aux_d.numel() * aux_d.element_size() if aux_d is not None else 0)
        link_total = float(link_gbs.item())
        e2e = {"value": cfg.pairs / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "host_input_bytes_per_step": host_bytes,
               "h2d_transport": ("register rows packed by the host threads inside the timed region (per row one base byte + 4-bit offsets "
                                 "+ <= 32 exceptions, csrc/hostpack.h) and unpacked on the device; a piece goes raw whenever the copy "
                                 "engine has run out of work; auxiliary sketches as they are" if packed_h2d else "raw register bytes"),
               "h2d_rows": load_info,
               "d2h_bytes_per_step": int(out_n * 16 + cfg.n * 8), "ms_per_step": e2e_ms, "steps": e2e_steps,
               "timer": "max(host wall clock, CUDA events) per step, max over ranks",
               "rank0_phases_ms": {k: v * 1e3 / e2e_steps for k, v in phase.items()},
               "h2d_link": {"measured_gbs_all_ranks": link_total, "achieved_gbs": h2d / (e2e_ms * 1e-3) / 1e9,
                            "frac": (h2d / (e2e_ms * 1e-3) / 1e9) / link_total if link_total > 0 else None,
                            "def": "pinned host -> device copy of the same register buffer alone, best of 3, summed over "
                                   "ranks copying at once; achieved = h2d bytes / whole e2e step"}}
        sel2.close()
        del regs_h, aux_h

    # ---- the other configurations of BASELINE.json, measured in the same run -------------------------
    configs = None
    if not a.no_configs:
        sel.close()
        del regs_d, aux_d
        torch.cuda.empty_cache()
        configs = {}
        for label, c in other_configs():
            r_d, a_d, kind = resident_inputs(c)
            s2 = S.Selection(local, stream=stream)
            s2.load(r_d, a_d, kind)
            if world > 1:
                sdist.setup_gather(s2)
            ms_c, acc_c, out_c, _ = timed_steps(s2, c, 3, max(1, a.cfg_steps), False)
            stc = acc_c[-1]
            kt = kernel_table(c, stc, mean_of(acc_c), sm_clk, n_sm, peak)
            d = max(kt, key=lambda e: e["ms"])
            entry = {"config": c.config(), "ms_per_step": ms_c, "value": c.pairs / (ms_c * 1e-3), "unit": UNIT,
                     "steps": max(1, a.cfg_steps), "warmup": 3,
                     "stages": {"pairs_cb": stc["pairs_cb"], "pairs_aux_rank0": stc["pairs_aux"], "pairs_out": out_c,
                                "bands_x_rows": [stc["n_bands"], stc["n_rows"]]},
                     "kernels_ms": {k: mean_of(acc_c)(f"ms_{k}") for k in ("bounds", "filter", "verify", "union", "estimate", "sort")}
                     | {"run_total": mean_of(acc_c)("ms_total")},
                     "dominant": {k: d[k] for k in ("stage", "kernel", "ms", "bound", "achieved", "peak", "unit", "frac")}}
            if world == 1 and rank == 0 and not a.no_config_parity and not a.no_cpu_baseline:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import oracle_api as O
                res = s2.run(tau=np.float32(c.tau), criterion=c.criterion)
                rh = r_d.cpu().numpy()
                ah = None if a_d is None else (a_d.cpu().numpy().view(np.uint64) if kind == AUX_SMH else a_d.cpu().numpy())
                ora = O.select(rh, 14, c.criterion, np.float32(c.tau), aux=ah, threads=os.cpu_count() or 1)
                p = diff_oracle(res, ora, c.tau)
                p["ok"] = parity_ok(p)
                entry["parity"] = {k: p[k] for k in ("ok", "pairs_oracle", "pairs_missing", "pairs_extra", "max_rel_jaccard",
                                                     "pairs_cb_equal", "pairs_aux_equal", "near_tau_count")}
                del res, rh, ah, ora
            configs[label] = entry
            s2.close()
            del r_d, a_d
            torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u8 registers / u64 buckets, f64 estimator", "data": "synthetic",
                "config": cfg.config(),
                "parallelism": (f"tile-shard x{world}" + (", peer-memory gather to rank 0; e2e: per-rank H2D slices + "
                                                           "NCCL all-gather" if world > 1 else "")),
                "stages": {"pairs_cb": st0["pairs_cb"], "pairs_aux_rank0": st0["pairs_aux"], "pairs_out": out_n,
                           "bands_x_rows": [st0["n_bands"], st0["n_rows"]]},
                "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(sum(s["launches"] for s in stats_acc)),
                "roofline": roofline, "kernels": kernels, "parity": parity, "cpu_baseline": cb, "cli": cli, "configs": configs}
        print(json.dumps(line), file=real_stdout, flush=True)
    if configs is None:
        sel.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
