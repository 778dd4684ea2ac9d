#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configurations, per the driver contract.

  python bench.py --gpus N --steps K --warmup W            this engine (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU path (oracle/_ref)

Metric: QPS at recall@10 >= 0.95 over the 17 filter fractions 2^-16..2^0 (k = 10, 10 000 queries per
fraction).  One "step" = one pass over all 17 fractions: for each fraction the batch of `nq` queries is answered
by the fastest (method, beam, final_multiply) operating point that reaches recall@10 >= 0.95 — methods:
prefilter (as the engine routes the batch by itself: one-launch kernel, task path or tensor-core sweep — same
rows), range-filter tree ("fenwick"), optimized postfilter (config 3: super optimized postfilter tree) — chosen
in an untimed sweep, exactly the pareto rule of the reference's plots (experiments/plot.py:14-28).
value = queries answered per second.

  value         inputs resident in HBM, CUDA-event timed on the engine's stream
  e2e           same step through the pybind classes the reference driver calls, ordinary pageable numpy arrays,
                default engine options (H2D of queries + windows and D2H of ids + dists inside the timed region)
  roofline      dominant kernel, algorithmic bytes / flops / CUDA-event kernel time
  cpu_baseline  the unmodified reference (oracle/_ref) on the box's host cores, bounded sample
  parity        the reference's rows against the engine's on the same queries / windows / graph files
                (prefilter rows equal up to 1e-5 ties; recall within 0.005 at equal method, beam, multiply)

Default (what the driver runs): --config c2 (BASELINE.json configs[1]).  Other configurations: c1, c3 (GloVe shape,
super tree), c4 (RedCaps shape; --rows scales it), c5 / c5adv (Deep shape / adversarial labels).
Multi-GPU: default = index replicated, every rank answers its own batch (weak scaling, no data-path collective);
--scaling strong = ONE batch cut into N slices; --mode label_shard = contiguous label ranges, in-library
ncclAllGather + merge; --mode group = ONE process, N GPUs behind the public batch_search call.  Time = max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[0] / configs[1]: the headline (c2 is what the driver's default run measures)
    "c1": dict(n=100_000, d=128, nq=10_000, seed=0, cutoff=1000, metric="l2", tree="wst",
               name="SIFT-shaped synthetic 100Kx128 fp32 L2"),
    "c2": dict(n=1_000_000, d=128, nq=10_000, seed=0, cutoff=1000, metric="l2", tree="wst",
               name="SIFT-shaped synthetic 1Mx128 fp32 L2"),
    # configs[2]: GloVe's size, angular (rows normalised, MIPS arithmetic), super optimized postfilter tree, swept over
    # the reference driver's beams x final_beam_multiplies (experiments/run_our_method.py:29-39,487-532)
    "c3": dict(n=1_183_514, d=100, nq=10_000, seed=0, cutoff=1000, metric="mips", angular=True, tree="super",
               beams=[10, 20, 40, 80, 160, 320, 640, 1280], mults=[1, 2, 3, 4, 8, 16, 32],
               name="GloVe-shaped synthetic 1.18Mx100 angular"),
    # configs[3]: RedCaps shape (CLIP-like 512-d angular, timestamp-style labels with many duplicates,
    # generate_redcaps_data.py:77-80), optimized postfilter; --n scales the row count (stated in the line)
    "c4": dict(n=12_000_000, d=512, nq=10_000, seed=0, cutoff=1000, metric="mips", angular=True, labels="timestamp_unique",
               tree="wst", name="RedCaps-shaped synthetic 12Mx512 angular, timestamp-style labels"),
    # configs[4]: Deep shape, label-range sharded (--mode label_shard / group); --n scales the row count
    "c5": dict(n=100_000_000, d=96, nq=10_000, seed=0, cutoff=1000, metric="l2", tree="wst",
               name="Deep-shaped synthetic 100Mx96 fp32 L2"),
    # configs[4], second half: the reference's adversarial recipe (generate_advserial_dataset.py:8-60: 100 clusters,
    # label ranges = clusters, query from cluster a with cluster b's range) plus windows whose smallest containing
    # bucket is the root (blow-up 2^4 ... 2^10)
    "c5adv": dict(n=1_000_000, d=96, nq=9_900, seed=0, cutoff=1000, metric="l2", tree="wst", adversarial=True,
                  name="adversarial 100-cluster synthetic 1Mx96 (normalised rows, L2)"),
}
ALL_POWERS = list(range(-16, 1))
POWERS = ALL_POWERS  # the fractions of the run (narrowed by --powers; "adv" / "blowup-p" keys for c5adv)
K = 10
BEAMS = [10, 20, 40, 80, 160, 320]
MULTS = [1, 2, 4]


def fractions_label() -> str:
    if POWERS == ALL_POWERS:
        return "2^-16..2^0"
    return ", ".join(frac_name(p) for p in POWERS)


def frac_name(p) -> str:
    return f"2^{p}" if isinstance(p, int) else str(p)


def make_inputs(cfg: dict, rank: int = 0, window_seed_shift: int = 0):
    """data, queries, labels, {fraction: windows[nq,2]} — the same on every rank except rank > 0's queries and
    windows of the weak-scaling run (own streams)."""
    from rangefilteredann_b200 import synth
    angular = bool(cfg.get("angular"))
    if cfg.get("adversarial"):
        data, queries, labels, adv_w = synth.make_adversarial(cfg["n"], cfg["d"], cfg["seed"])
        windows = {}
        for p in POWERS:
            if p == "adv":
                windows[p] = adv_w
            else:
                windows[p] = synth.make_blowup_windows(labels, int(str(p).split("blowup")[1]), len(queries), seed=1000 + window_seed_shift)
        return data, queries, labels, windows
    data, queries, labels = synth.make_dataset(cfg["n"], cfg["d"], cfg["nq"], cfg["seed"], angular=angular,
                                               label_kind=cfg.get("labels", "unique"))
    if rank > 0:
        queries = synth.make_rank_queries(cfg["d"], cfg["nq"], cfg["seed"], rank, angular=angular)
    windows = {p: synth.make_windows(labels, p, cfg["nq"], seed=1000 + window_seed_shift + p) for p in POWERS}
    return data, queries, labels, windows


def class_suffix(cfg: dict) -> str:
    return "FloatMips" if cfg["metric"] == "mips" else "FloatEuclidian"


def make_indices(mod, cfg: dict, cdir: str, data, labels):
    """(tree index, prefilter index) of `mod` (this engine or the reference — same class names and arguments)."""
    sfx = class_suffix(cfg)
    bp = mod.BuildParams(64, 500, 1.0, cdir)
    if cfg["tree"] == "super":
        tree = getattr(mod, "SuperOptimizedPostfilterTreeIndex" + sfx)(data, labels, cfg["cutoff"], 2.0, 0.5, bp)
    else:
        tree = getattr(mod, "VamanaRangeFilterTreeIndex" + sfx)(data, labels, cfg["cutoff"], 2, bp)
    pre = getattr(mod, "PrefilterIndex" + sfx)(data, labels)
    return tree, pre


def tree_methods(cfg: dict):
    return ("super",) if cfg["tree"] == "super" else ("fenwick", "optimized_postfilter")
RECALL_TARGET = 0.95
PREFILTER_OPS = ("prefilter", "prefilter_direct", "prefilter_tc")
# the engine's own routing of a PrefilterIndex batch (csrc/wsann.cu ws_run_batch, defaults of the options
# gemm_min_window / scan_chunk): what `PrefilterIndex*.batch_search` does when nobody sets an option
GEMM_MIN_WINDOW = 768
SCAN_CHUNK = 8192


def auto_prefilter_route(mean_window: float) -> str:
    if mean_window >= GEMM_MIN_WINDOW:
        return "prefilter_tc"
    if mean_window <= SCAN_CHUNK:
        return "prefilter_direct"
    return "prefilter"


def workload_string(cfg: dict) -> str:
    """One string for both arms (the driver compares them)."""
    labels = {"timestamp": "timestamp-style integer labels (duplicates)",
              "timestamp_unique": "timestamp-style labels (integer seconds as float32, ties nudged apart by ulps: the reference's "
                                  "label sort is unstable, so one shared set of graph files needs a total label order)"}.get(
        cfg.get("labels", ""), "uniform unique labels")
    if cfg.get("adversarial"):
        labels = "cluster-aligned labels"
    if cfg["tree"] == "super":
        tree = ("super optimized postfilter tree (cutoff 1000, split 2, shift 0.5, R=64 L=500 alpha=1, one set of "
                "reference-format graph files searched by both arms), prefilter / super optimized postfilter")
    else:
        tree = ("2-WST (cutoff 1000, R=64 L=500 alpha=1, one set of reference-format graph files searched by both arms), "
                "prefilter / range-filter tree / optimized postfilter")
    scaled = f" [rows scaled to {cfg['n']} for this run]" if cfg.get("scaled") else ""
    return (f"{cfg['name']}{scaled}, {labels}, {tree}, {len(POWERS)} fractions x {cfg['nq']} queries, k=10, "
            f"per fraction the fastest method reaching recall@10 >= 0.95")


def rows_equal_up_to_ties(ids, dists, rids, rdists, rtol=1e-5, scale=0.0):
    """Per row: ids equal position-wise, except among entries whose distances tie within rtol (BASELINE.json:
    'prefilter top-k ids must match the reference exactly, except for distance ties within 1e-5 relative').
    `scale`: magnitude the tolerance is relative to when a distance is a difference of larger terms — an inner
    product of unit vectors is a sum of d products of magnitude |q||x| = 1 that may cancel to ~0, so MIPS rows are
    compared within rtol * max(|dist|, 1)."""
    ok = np.zeros(len(ids), dtype=bool)
    for i in range(len(ids)):
        if not np.allclose(dists[i], rdists[i], rtol=rtol, atol=max(1e-30, rtol * scale)):
            continue
        if np.array_equal(ids[i], rids[i]):
            ok[i] = True
            continue
        good = True
        for j in np.nonzero(ids[i] != rids[i])[0]:
            d = rdists[i, j]
            tie = np.isclose(rdists[i], d, rtol=rtol, atol=1e-30)
            if not (ids[i, j] in rids[i][tie] or np.isclose(rdists[i, -1], d, rtol=rtol, atol=1e-30)):
                good = False
                break
        ok[i] = good
    return ok


# oracle/_ref is the reference compiled with its own flags (CMakeLists.txt:17-24) except that -march=native
# becomes -march=x86-64-v3 (AVX2 + FMA; the build container's CPU is not the GPU box's)
REF_MARCH = "x86-64-v3 (oracle/Makefile; the reference's CMake uses -march=native)"


def reduce_max(values, device=None) -> list[float]:
    """max over ranks of a small vector of timings (all ranks get the result)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_rows(local: np.ndarray, world: int):
    """Concatenate per-rank result rows on rank 0 (strong scaling: the host gather of nq x k ids; the data path
    itself needs no collective)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return local
    objs = [None] * world if dist.get_rank() == 0 else None
    dist.gather_object(local, objs, dst=0)
    return np.concatenate(objs, axis=0) if dist.get_rank() == 0 else None


def broadcast_object(obj, src: int = 0):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def broadcast_bytes(payload: bytes | None, src: int = 0) -> bytes:
    """Rank `src` hands a small byte string (the NCCL id of the in-library communicator) to every rank."""
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


_RESULT_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner on stdout at init, the reference prints one line per loaded graph), so file descriptor 1 is
    pointed at stderr for the whole run and the result line goes to a private copy of the original."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------
# data, graphs, ground truth
# ------------------------------------------------------------------------------------------
def cache_dir(cfg_name: str) -> str:
    cfg = CONFIGS[cfg_name]
    tag = cfg_name + (f"-n{cfg['n']}" if cfg.get("scaled") else "")
    return os.path.join(ROOT, "data_cache", tag, cfg["tree"]) + "/"


def expected_graph_count(n: int, cutoff: int, split: int = 2) -> int:
    rows, size = 1, n
    total, nb = 1, 1
    while size > cutoff:
        size = (size + split - 1) // split
        nb *= split
        total += nb
        rows += 1
    return total


def load_ref():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import _load_ext, find_ext
    path = find_ext(os.path.join(ROOT, "oracle", "_ref"))
    if path is None:
        return None
    os.environ.setdefault("PARLAY_NUM_THREADS", str(os.cpu_count()))
    return _load_ext(path)


def validate_cache(cdir: str, data, labels):
    """Graph files are named after (L, R, alpha, min label, max label, count) only
    (postfilter_vamana.h:126-132), so a cache left by a different dataset with the same label VALUES
    would be loaded silently.  A fingerprint of the points and labels guards the directory: on a
    mismatch (or a cache of unknown origin) the files are dropped and rebuilt."""
    import hashlib
    os.makedirs(cdir, exist_ok=True)
    h = hashlib.sha1()
    h.update(np.ascontiguousarray(labels).tobytes())
    h.update(np.ascontiguousarray(data[::1009]).tobytes())
    fp = h.hexdigest()
    path = os.path.join(cdir, "FINGERPRINT")
    old = open(path).read().strip() if os.path.exists(path) else None
    bins = [f for f in os.listdir(cdir) if f.endswith(".bin")]
    if old != fp and bins:
        log(f"graph cache {cdir} belongs to another dataset (fingerprint {old} != {fp}): dropping {len(bins)} files")
        for f in bins:
            os.remove(os.path.join(cdir, f))
    if old != fp:
        with open(path, "w") as f:
            f.write(fp + "\n")


def ensure_graphs(cfg_name: str, cfg: dict, data, labels, rank: int):
    """Both arms search the same reference-format graph files under data_cache/<cfg>/wst/.
    A missing cache is produced once, untimed: by this engine's device-side builder when a
    GPU is visible (the index constructor builds and saves whatever is missing), else by the
    reference builder (oracle/_ref)."""
    cdir = cache_dir(cfg_name)
    want = expected_graph_count(cfg["n"], cfg["cutoff"]) if cfg["tree"] == "wst" else 1
    have = len([f for f in os.listdir(cdir) if f.endswith(".bin")]) if os.path.isdir(cdir) else 0
    if (have >= want and cfg["tree"] == "wst") or rank != 0:
        return cdir
    os.makedirs(cdir, exist_ok=True)
    t0 = time.time()
    from rangefilteredann_b200 import load_engine
    eng = load_engine()
    if eng.device_count() > 0:
        log(f"graph cache {cdir} has {have} files: loading / building the rest on the GPU (untimed setup)")
        make_indices(eng, cfg, cdir, data, labels)
    else:
        log(f"graph cache {cdir} has {have} files: building with the reference builder (untimed setup)")
        ref = load_ref()
        if ref is None:
            raise SystemExit("no graph cache, no GPU and no oracle/_ref to build it with")
        make_indices(ref, cfg, cdir, data, labels)
    log(f"graph cache built in {time.time() - t0:.1f}s")
    return cdir


def ground_truth_torch(data, queries, labels, windows_by_power, device, metric="l2"):
    """Closed-interval brute-force top-10 (filter_generation_utils.py:142-168) in fp32 on
    the GPU with torch — independent of the engine under test."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    X = torch.from_numpy(data).to(device)
    L = torch.from_numpy(labels).to(device)
    Q = torch.from_numpy(queries).to(device)
    xn = (X * X).sum(1)
    out = {}
    chunk = max(1, min(len(queries), (1 << 30) // max(1, len(data))))
    for p, w in windows_by_power.items():
        W = torch.from_numpy(w).to(device)
        nq_p = len(w)
        gt = np.empty((nq_p, K), np.int64)
        for s in range(0, nq_p, chunk):
            e = min(nq_p, s + chunk)
            d = (xn[None, :] - 2.0 * (Q[s:e] @ X.T)) if metric == "l2" else -(Q[s:e] @ X.T)
            mask = (L[None, :] >= W[s:e, 0:1]) & (L[None, :] <= W[s:e, 1:2])
            d = torch.where(mask, d, torch.full_like(d, float("inf")))
            vals, idx = torch.topk(d, K, dim=1, largest=False)
            idx = torch.where(torch.isinf(vals), torch.full_like(idx, -1), idx)
            gt[s:e] = idx.cpu().numpy()
        out[p] = gt
    del X, Q, L
    torch.cuda.empty_cache()
    return out


def recall_at_k(ids: np.ndarray, gt: np.ndarray) -> float:
    """mean |top-k ∩ GT| / |GT| (GT pads -1 ignored; result pads never match)."""
    valid = gt >= 0
    hit = ((gt[:, :, None] == ids[:, None, :].astype(np.int64)).any(2) & valid).sum(1)
    denom = np.maximum(valid.sum(1), 1)
    return float(np.mean(hit / denom))


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML in a thread (a sample every
    ~2 ms, the timed region of the default run is tens of milliseconds — too short for `nvidia-smi -lms`,
    which needs longer than that to start); `nvidia-smi` only when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.samples = []   # (sm_mhz, reasons bitmask)
        self.stop_flag = False
        self.thread = None
        self.nvml = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES if it remaps them
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(mhz), int(reasons)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            nv = self.nvml
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [m for m, _ in self.samples]
            seen = 0
            for _, r in self.samples:
                seen |= r
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(name for name, b in bits.items() if seen & b), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------
# the engine arm
# ------------------------------------------------------------------------------------------
class EngineRunner:
    """Device-resident and host-buffer execution of one (fraction, operating point)."""

    def __init__(self, tree_index, nq: int, d: int, cfg: dict):
        from rangefilteredann_b200 import capi
        self.capi = capi
        self.tree = tree_index
        self.cfg = cfg
        self.h = capi.Handle.borrow(tree_index)
        self.nq, self.d = nq, d
        self.dq = self.h.dalloc(nq * d * 4)
        self.dids = self.h.dalloc(nq * K * 4)
        self.ddists = self.h.dalloc(nq * K * 4)
        self.dwin = {}
        self.mean_window = {}

    def upload(self, queries, windows_by_power, sorted_labels=None):
        self.h.h2d(self.dq, queries)
        for p, w in windows_by_power.items():
            self.dwin[p] = self.h.dalloc(w.nbytes)
            self.h.h2d(self.dwin[p], w)
            if sorted_labels is not None:
                rows = np.searchsorted(sorted_labels, w[:, 1]) - np.searchsorted(sorted_labels, w[:, 0])
                self.mean_window[p] = float(np.mean(rows.clip(min=0)))

    def launch_dev(self, power, op):
        method, beam, mult = op
        if method in PREFILTER_OPS:
            # three ways the engine answers the same PrefilterIndex::batch_search, rows bit-identical:
            # "prefilter" = task path (K3 -> K1 scan -> K4), "prefilter_direct" = the one-launch kernel for
            # small windows (K1d), "prefilter_tc" = tensor-core sweep + fp32 re-rank (csrc/ws_gemm.cu)
            self.h.set_option("gemm_prefilter", 1 if method == "prefilter_tc" else 0)
            self.h.set_option("prefilter_direct", 1 if method == "prefilter_direct" else 0)
            self.h.prefilter_batch(self.dq, self.dwin[power], self.nq, K, self.dids, self.ddists, device_ptrs=True)
        else:
            qp = self.capi.query_params(k=K, beam=beam, final_multiply=mult)
            self.h.tree_batch(method, self.dq, self.dwin[power], self.nq, qp, self.dids, self.ddists, device_ptrs=True)

    def sync(self):
        self.h.sync()

    def fetch(self):
        ids = np.empty((self.nq, K), np.uint32)
        self.h.d2h(ids, self.dids)
        return ids

    def time_dev(self, power, op, reps=2):
        best = 1e30
        for _ in range(reps):
            self.h.timer_start()
            self.launch_dev(power, op)
            best = min(best, self.h.timer_stop())
        return best


def choose_operating_points(runner, gts, rank, cfg):
    """Untimed sweep: per fraction and method, the first (smallest) beam reaching the recall
    target for each final_multiply; the fastest of those is the method's operating point."""
    table = {}
    beams = cfg.get("beams", BEAMS)
    mults = cfg.get("mults", MULTS)
    for p in POWERS:
        per_method = {}
        # prefilter: exact.  Slow variants are not timed where they are hopeless (the one-launch kernel and the
        # task path re-read every window once per query: seconds per batch on multi-million-row windows)
        auto = auto_prefilter_route(runner.mean_window[p])
        heavy = runner.mean_window[p] * runner.nq * cfg["d"] * 4 > 4e12
        for alt in PREFILTER_OPS:
            if alt not in getattr(runner, "prefilter_ops", PREFILTER_OPS) or (heavy and alt != auto):
                continue
            runner.launch_dev(p, (alt, 0, 0))
            r = recall_at_k(runner.fetch(), gts[p])
            ms = runner.time_dev(p, (alt, 0, 0))
            per_method[alt] = dict(op=(alt, 0, 0), recall=r, ms=ms)
        # what PrefilterIndex.batch_search picks by itself for this fraction's windows (the step uses THIS, not the
        # fastest of the three)
        if auto in per_method:
            per_method["prefilter_auto"] = dict(per_method[auto])
        for method in tree_methods(cfg):
            best = None
            for mult in ([1] if method == "fenwick" else mults):
                for beam in beams:
                    op = (method, beam, mult)
                    runner.launch_dev(p, op)
                    r = recall_at_k(runner.fetch(), gts[p])
                    if r >= RECALL_TARGET:
                        ms = runner.time_dev(p, op)
                        if best is None or ms < best["ms"]:
                            best = dict(op=op, recall=r, ms=ms)
                        # fenwick gets slower with the beam; optimized postfilter need not (a
                        # larger start beam can save doubling rounds), so keep looking there
                        if method == "fenwick" or (best is not None and ms > 2.0 * best["ms"]):
                            break
            if best is not None:
                per_method[method] = best
        table[p] = per_method
        if rank == 0:
            log(f"{frac_name(p)}: " + ", ".join(f"{m}: beam {v['op'][1]} x{v['op'][2]} recall {v['recall']:.4f} {v['ms']:.3f} ms"
                                        for m, v in per_method.items()))
    return table


def run_engine(args, rank, world, local_rank):
    from rangefilteredann_b200 import capi, load_engine, synth
    eng = load_engine()
    if eng.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device visible — this engine has no CPU fallback")
    cfg = CONFIGS[args.config]
    os.environ["WSANN_DEVICE"] = str(local_rank)
    t_setup = time.time()
    # data and labels are the same on every rank, for every world size and in the reference arm (the graph
    # cache is keyed by them).  Weak scaling: each rank answers its own batch of queries.  Strong scaling: the ONE
    # batch of rank 0 is cut into contiguous slices, one per rank (no data-path collective).
    from rangefilteredann_b200 import sharding
    strong = args.scaling == "strong" and world > 1
    data, queries, labels, windows = make_inputs(cfg, 0 if strong else rank, 0 if strong else 17 * rank)
    nq_rank = cfg["nq"]
    if strong:
        lo, hi = sharding.shard_bounds(cfg["nq"], rank, world)
        queries = np.ascontiguousarray(queries[lo:hi])
        windows = {p: np.ascontiguousarray(w[lo:hi]) for p, w in windows.items()}
        nq_rank = hi - lo
    cdir = cache_dir(args.config)
    os.makedirs(cdir, exist_ok=True)
    if rank == 0:
        validate_cache(cdir, data, labels)
    if world > 1:  # rank 0 fills the cache (its constructor builds + saves what is missing), the others load it
        import torch.distributed as dist
        if rank == 0:
            ensure_graphs(args.config, cfg, data, labels, 0)
        dist.barrier()
    # the classes run_our_method.py:249,357,387,516 call (tree methods; "prefiltering")
    tree, pre = make_indices(eng, cfg, cdir, data, labels)
    log(f"rank {rank}: data + index ready in {time.time() - t_setup:.1f}s")
    gts = ground_truth_torch(data, queries, labels, windows, f"cuda:{local_rank}", cfg["metric"])
    sorted_labels = np.sort(labels)
    runner = EngineRunner(tree, nq_rank, cfg["d"], cfg)
    runner.upload(queries, windows, sorted_labels)
    h = runner.h
    if args.ops_file and os.path.exists(args.ops_file):
        saved = json.load(open(args.ops_file))
        table = {(int(p) if p.lstrip("-").isdigit() else p): {m: dict(op=tuple(v["op"]), recall=v["recall"], ms=v["ms"]) for m, v in pm.items()}
                 for p, pm in saved.items()}
    else:
        table = choose_operating_points(runner, gts, rank, cfg)
        if args.ops_file and rank == 0:
            json.dump({str(p): {m: dict(op=list(v["op"]), recall=v["recall"], ms=v["ms"]) for m, v in pm.items()}
                       for p, pm in table.items()}, open(args.ops_file, "w"))
    # per fraction: the fastest of {prefilter as the engine routes it by itself, range-filter tree, optimized postfilter}
    step_methods = ("prefilter_auto",) + tree_methods(cfg)
    ops = {p: min((table[p][m] for m in step_methods if m in table[p]), key=lambda v: v["ms"])["op"] for p in POWERS}
    if strong:  # every rank must run the same operating points: rank 0's choice
        ops = broadcast_object(ops)
    nq_step = nq_rank * len(POWERS)
    nq_step_total = cfg["nq"] * len(POWERS) if strong else nq_rank * len(POWERS) * world

    def barrier():
        h.sync()
        if world > 1:
            import torch
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def step_dev():
        for p in POWERS:
            runner.launch_dev(p, ops[p])

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        step_dev()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    l0 = h.launches()
    h.timer_start()
    for _ in range(args.steps):
        step_dev()
    ms_total = h.timer_stop()
    barrier()
    launches = h.launches() - l0
    clk = clocks.stop()

    # ---- per-kernel times + counters over the same K steps (roofline)
    h.set_option("profile_kernels", 1)
    h.reset_stats()
    h.kernel_times(reset=True)
    for _ in range(args.steps):
        step_dev()
    ktimes = h.kernel_times(reset=True)
    stats = h.stats()
    h.set_option("profile_kernels", 0)

    # ---- end to end through the public API: the pybind classes the reference driver calls (run_our_method.py:249,357,
    # 387), ordinary pageable numpy arrays, default options (the engine routes prefilter batches by itself)
    h.set_option("gemm_prefilter", 2)
    h.set_option("prefilter_direct", 2)
    hpre = capi.Handle.borrow(pre)
    qps_by_op = {}

    def e2e_call(p):
        method, beam, mult = ops[p]
        if method in PREFILTER_OPS:
            return pre.batch_search(queries, windows[p], nq_rank, qps_by_op.setdefault((10, 1), eng.QueryParams(K, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False)))
        qp = qps_by_op.setdefault((beam, mult), eng.QueryParams(K, beam, 1.35, 10_000_000, 10_000, mult, 10000, None, False))
        if method == "super":
            return tree.batch_search(queries, windows[p], nq_rank, qp)
        return tree.batch_search(queries, windows[p], nq_rank, method, qp)

    def step_e2e():
        for p in POWERS:
            ids, _ = e2e_call(p)
        return ids

    # which kernels the auto-routed prefilter calls actually launch (must be what the device-timed step runs)
    routes = {}
    hpre.set_option("profile_kernels", 1)
    for p in POWERS:
        if ops[p][0] in PREFILTER_OPS:
            hpre.kernel_times(reset=True)
            e2e_call(p)
            kt = hpre.kernel_times(reset=True)
            routes[frac_name(p)] = "prefilter_tc" if "gemm_sweep" in kt else ("prefilter" if "decompose" in kt else "prefilter_direct")
            if routes[frac_name(p)] != ops[p][0]:
                log(f"WARNING {frac_name(p)}: auto routing took {routes[frac_name(p)]}, the device-timed step ran {ops[p][0]}")
    hpre.set_option("profile_kernels", 0)
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    h.sync()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- reduce over ranks (max time)
    ms_total, e2e_ms = reduce_max([ms_total, e2e_s * 1000.0], device=f"cuda:{local_rank}" if world > 1 else None)
    if rank != 0:
        return

    ms_per_step = ms_total / args.steps
    value = nq_step_total / (ms_per_step / 1000.0)
    e2e_value = nq_step_total / (e2e_ms / args.steps / 1000.0)

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    dpad_bytes = ((cfg["d"] * 4 + 63) // 64) * 64
    nq_step = nq_rank * len(POWERS)
    beam_ms = sum(v["ms"] for kname, v in ktimes.items() if kname.startswith("beam"))
    beam_launches = sum(v["launches"] for kname, v in ktimes.items() if kname.startswith("beam"))
    scan_ms = ktimes.get("scan", {}).get("ms", 0.0)
    # graph search: visited * R*4 + dist_cmps * d_pad*4 + (beam*4 per search)
    beam_bytes = stats["visited"] * 64 * 4 + stats["dist_cmps"] * dpad_bytes + stats["beam_sum"] * 4
    # tensor-core prefilter sweep: 2*d flops per (query, in-window point); the points are
    # counted from the windows of the fractions whose operating point is prefilter_tc
    gemm_ms = ktimes.get("gemm_sweep", {}).get("ms", 0.0)
    gemm_pairs = sum(int(np.sum((np.searchsorted(np.sort(labels), windows[p][:, 1]) -
                                 np.searchsorted(np.sort(labels), windows[p][:, 0])).clip(min=0)))
                     for p in POWERS if ops[p][0] == "prefilter_tc") * args.steps
    gemm_flops = 2.0 * cfg["d"] * gemm_pairs
    tensor_peak = float(peaks.get("bf16_tflops", 1590.0))
    gemm_info = None
    if gemm_ms > 0:
        gemm_info = {"kernel": "ws_gemm_topk_kernel (tcgen05 kind::f16, fp32 accumulate, fp32 re-rank)", "bound": "tensor",
                     "achieved": round(gemm_flops / (gemm_ms / 1000.0) / 1e12, 1), "peak": tensor_peak, "unit": "TFLOP/s",
                     "frac": round(gemm_flops / (gemm_ms / 1000.0) / 1e12 / tensor_peak, 4),
                     "peak_source": "measured dense bf16 (MEASURED_PEAKS.json bf16_tflops); flops = 2*d per (query, "
                                    "in-window point) — tiles also score the out-of-window points of a group's union",
                     "ms_per_step": round(gemm_ms / args.steps, 4),
                     "flops_per_step": gemm_flops / args.steps}
    # streaming-scan prefilter (task path or one-launch kernel): rows * d_pad * 4, rows counted
    # from the windows of the fractions answered that way (the device counter also counts the tensor-core
    # path's windows, so it cannot be used here)
    sorted_labels = np.sort(labels)
    win_rows = {p: int(np.sum((np.searchsorted(sorted_labels, windows[p][:, 1]) -
                               np.searchsorted(sorted_labels, windows[p][:, 0])).clip(min=0))) for p in POWERS}
    scan_bytes = sum(win_rows[p] for p in POWERS if ops[p][0] in ("prefilter", "prefilter_direct")) * args.steps * dpad_bytes
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        tr = {}

    def traffic_of(key):
        ent = tr.get(key)
        if not ent:
            return None, None
        return int(ent["traffic_bytes"]), (f"ncu --set full capture of one launch ({ent['launch']}): {ent['traffic_bytes'] / 1e9:.2f} GB "
                                           f"DRAM read+write for {ent['algorithmic']} algorithmic ({ent['source']}); the per-launch "
                                           f"figures beside it are means over this run's launches")

    cands = []
    if beam_ms > 0:
        # per-launch figures refer to the launches of the tier that carries the time (each batch also launches the
        # tail tiers, which find their queues empty and return within microseconds)
        top = max((kn for kn in ktimes if kn.startswith("beam")), key=lambda kn: ktimes[kn]["ms"])
        nl = max(1, ktimes[top]["launches"])
        ach = beam_bytes / (beam_ms / 1000.0) / 1e9
        t, tn = traffic_of("ws_beam_warp_kernel")
        cands.append({"bound": "hbm", "kernel": "ws_beam_warp_kernel (+ ws_beam_cta2_kernel tail tiers)", "achieved": round(ach, 1),
                      "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": t, "traffic_note": tn,
                      "peak_source": peak_src, "bytes_per_launch": int(beam_bytes / nl), "ms_per_launch": round(beam_ms / nl, 4),
                      "ms_per_step": round(beam_ms / args.steps, 4)})
    if scan_ms > 0 and scan_bytes > 0:
        nl = max(1, ktimes["scan"]["launches"])
        ach = scan_bytes / (scan_ms / 1000.0) / 1e9
        t, tn = traffic_of("ws_scan_warp_kernel")
        cands.append({"bound": "hbm", "kernel": "ws_scan_warp_kernel / ws_prefilter_direct_kernel", "achieved": round(ach, 1),
                      "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": t, "traffic_note": tn,
                      "peak_source": peak_src + "; windows of one batch overlap, so rows are shared through L2 and the "
                                                "no-reuse byte count can exceed the HBM peak (SURVEY.md §8d)",
                      "bytes_per_launch": int(scan_bytes / nl), "ms_per_launch": round(scan_ms / nl, 4),
                      "ms_per_step": round(scan_ms / args.steps, 4)})
    if gemm_info:
        nl = max(1, ktimes["gemm_sweep"]["launches"])
        t, tn = traffic_of("ws_gemm_topk_kernel")
        g = dict(gemm_info)
        g.update({"traffic": t, "traffic_note": tn, "flops_per_launch": gemm_flops / nl, "ms_per_launch": round(gemm_ms / nl, 4)})
        cands.append(g)
    cands.sort(key=lambda c: -c["ms_per_step"])
    roofline = dict(cands[0]) if cands else {"bound": "hbm", "achieved": 0.0, "peak": peak, "unit": "GB/s", "frac": 0.0, "traffic": None}
    roofline["dominant_by"] = "kernel time per step (CUDA events)"
    roofline["other_kernels"] = cands[1:]
    roofline["kernel_ms_per_step"] = {kname: round(v["ms"] / args.steps, 4) for kname, v in ktimes.items()}

    cpu, parity = (None, None)
    if not args.no_cpu:
        cpu, parity = cpu_baseline(args, cfg, data, queries, labels, windows, gts, table, ops, eng, tree, pre)

    per_fraction = {}
    for p in POWERS:
        per_fraction[frac_name(p)] = {m: {"beam": v["op"][1], "final_multiply": v["op"][2], "recall": round(v["recall"], 4),
                                          "qps": round(nq_rank / (v["ms"] / 1000.0))} for m, v in table[p].items()}
        per_fraction[frac_name(p)]["best"] = ops[p][0]
    line = {
        "metric": "QPS at recall@10>=0.95, one pass over filter fractions " + fractions_label(),
        "value": round(value, 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(cfg),
                   "name": args.config, "queries_per_step": nq_step_total, "queries_per_rank_per_fraction": nq_rank,
                   "l2_policy": "working set (vectors + adjacency, >= 0.25 GB, random gathers) exceeds the 126 MB L2; no flush",
                   "parallelism": f"query-sharded dp{world}, index replicated"},
        "e2e": {"value": round(e2e_value, 1), "unit": "queries/s",
                "h2d_bytes_per_step": int(len(POWERS) * nq_rank * (cfg["d"] * 4 + 8)),
                "d2h_bytes_per_step": int(len(POWERS) * nq_rank * K * 8),
                "api": f"pybind PrefilterIndex{class_suffix(cfg)}.batch_search / {type(tree).__name__}.batch_search, "
                       "pageable numpy arrays, default engine options (prefilter batches routed by the engine)",
                "prefilter_routes": routes},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
        "counters_per_step": {kname: int(v / args.steps) for kname, v in stats.items()},
        "per_fraction": per_fraction,
    }
    if args.results_csv:
        # the sweep in the reference driver's own results format (experiments/run_our_method.py:538-567),
        # readable by its experiments/plot.py
        from rangefilteredann_b200 import results as res
        rows = [(res.filter_width_name(p), res.method_name(m, v["op"][1], max(1, v["op"][2])), v["recall"], v["ms"] / 1000.0)
                for p in POWERS for m, v in table[p].items()]
        res.save_results(rows, args.results_csv, cfg["nq"], f"B200x{world}")
    emit(line)


# ------------------------------------------------------------------------------------------
# label-range sharded mode (one process per GPU) and the in-engine group mode (one process, N GPUs)
# ------------------------------------------------------------------------------------------
class ShardRunner:
    """One rank of the label-sharded run: local search on the shard + ncclAllGather + merge, all enqueued on the
    index stream by libwsann_cuda.so (rangefilteredann_b200/label_shard.py).  Timings are max over ranks so that every
    rank takes the same sweep decisions (the calls are collective)."""

    def __init__(self, shard, nq: int, cfg: dict, world: int, device):
        from rangefilteredann_b200 import capi
        self.capi, self.shard, self.nq, self.cfg, self.world, self.device = capi, shard, nq, cfg, world, device
        self.h = shard.h
        self.mean_window = {}
        self.windows = {}
        self.queries = None
        self.cur = None

    def upload(self, queries, windows_by_power, sorted_labels):
        self.queries = queries
        for p, w in windows_by_power.items():
            self.windows[p] = w
            rows = np.searchsorted(sorted_labels, w[:, 1]) - np.searchsorted(sorted_labels, w[:, 0])
            self.mean_window[p] = float(np.mean(rows.clip(min=0))) / self.world  # per shard

    def _stage(self, power):
        if self.cur != power:
            self.shard.upload(self.queries, self.windows[power], K)
            self.cur = power

    def launch_dev(self, power, op):
        method, beam, mult = op
        self._stage(power)
        if method in PREFILTER_OPS:
            self.h.set_option("gemm_prefilter", 1 if method == "prefilter_tc" else 0)
            self.h.set_option("prefilter_direct", 1 if method == "prefilter_direct" else 0)
            self.shard.search_device(self.nq, "prefilter", None, K)
        else:
            self.shard.search_device(self.nq, method, self.capi.query_params(k=K, beam=beam, final_multiply=mult), K)

    def sync(self):
        self.h.sync()

    def fetch(self):
        return self.shard.fetch(self.nq, K)[0]

    def time_dev(self, power, op, reps=2):
        best = 1e30
        for _ in range(reps):
            self.h.timer_start()
            self.launch_dev(power, op)
            best = min(best, self.h.timer_stop())
        return reduce_max([best], device=self.device)[0]


def run_label_shard(args, rank, world, local_rank):
    """BASELINE.json configs[4]: every rank owns a contiguous label range of the data set with its own B-WST (graphs
    built on its GPU), every rank answers the WHOLE batch, the partial [nq][k] rows are all-gathered over NVLink and
    merged per query inside the library.  value = queries of the one shared batch per second (max over ranks)."""
    from rangefilteredann_b200 import capi, label_shard, load_engine
    eng = load_engine()
    if eng.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device visible — this engine has no CPU fallback")
    cfg = CONFIGS[args.config]
    os.environ["WSANN_DEVICE"] = str(local_rank)
    dev = f"cuda:{local_rank}" if world > 1 else None
    t_setup = time.time()
    data, queries, labels, windows = make_inputs(cfg)
    nq = len(queries)
    uid = None
    if world > 1:
        uid = broadcast_bytes(capi.nccl_unique_id() if rank == 0 else None)
    tag = args.config + (f"-n{cfg['n']}" if cfg.get("scaled") else "")
    shard = label_shard.LabelShardedTree(data, labels, rank, world, os.path.join(ROOT, "data_cache", tag, "label_shards"),
                                         cutoff=cfg["cutoff"], metric=cfg["metric"], unique_id=uid)
    log(f"rank {rank}: shard of {len(shard.owned)} points + tree ready in {time.time() - t_setup:.1f}s")
    gts = ground_truth_torch(data, queries, labels, windows, f"cuda:{local_rank}", cfg["metric"])
    runner = ShardRunner(shard, nq, cfg, world, dev)
    runner.upload(queries, windows, np.sort(labels))
    h = shard.h
    table = choose_operating_points(runner, gts, rank, cfg)
    step_methods = PREFILTER_OPS + tree_methods(cfg)
    ops = {p: min((table[p][m] for m in step_methods if m in table[p]), key=lambda v: v["ms"])["op"] for p in POWERS}
    ops = broadcast_object(ops)

    def barrier():
        h.sync()
        if world > 1:
            import torch
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def step_dev():
        for p in POWERS:
            runner.launch_dev(p, ops[p])

    for _ in range(args.warmup):
        step_dev()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    l0 = h.launches()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_dev()
    h.sync()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    barrier()
    launches = h.launches() - l0
    clk = clocks.stop()
    # device time of the kernels + collectives alone (queries already resident): per fraction, CUDA events
    dev_ms, search_ms = 0.0, 0.0
    per_fraction = {}
    for p in POWERS:
        t_all = runner.time_dev(p, ops[p], reps=3)
        # the local search alone (communicator bypassed): what is left is the exchange + merge
        shard._has_comm, had = False, shard._has_comm
        t_search = runner.time_dev(p, ops[p], reps=3)
        shard._has_comm = had
        dev_ms += t_all
        search_ms += t_search
        per_fraction[frac_name(p)] = {"best": ops[p][0], "beam": ops[p][1], "final_multiply": ops[p][2], "recall": round(table[p][ops[p][0] if ops[p][0] in table[p] else "prefilter"]["recall"], 4),
                                      "ms": round(t_all, 4), "local_search_ms": round(t_search, 4),
                                      "exchange_merge_ms": round(max(0.0, t_all - t_search), 4),
                                      "qps": round(nq / (t_all / 1000.0))}
    # end to end: host numpy in, host numpy out, every step
    qps = {p: capi.query_params(k=K, beam=max(1, ops[p][1]), final_multiply=max(1, ops[p][2])) for p in POWERS}

    def step_e2e():
        for p in POWERS:
            m = ops[p][0]
            if m in PREFILTER_OPS:
                h.set_option("gemm_prefilter", 1 if m == "prefilter_tc" else 0)
                h.set_option("prefilter_direct", 1 if m == "prefilter_direct" else 0)
            shard.batch_search(queries, windows[p], "prefilter" if m in PREFILTER_OPS else m, qps[p], K)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    barrier()
    wall_ms, e2e_ms, dev_ms, search_ms = reduce_max([wall_ms, e2e_ms, dev_ms, search_ms], device=dev)
    if rank != 0:
        return
    nq_step = nq * len(POWERS)
    ms_per_step = wall_ms / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    line = {
        "metric": "QPS at recall@10>=0.95, one pass over filter fractions " + fractions_label(),
        "value": round(nq_step / (ms_per_step / 1000.0), 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "strong (label-sharded: the data set is cut into one label range per GPU, every GPU answers the whole batch)",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(cfg), "name": args.config, "mode": "label_shard", "rows_total": int(len(labels)),
                   "rows_per_gpu": int(len(shard.owned)), "queries_per_step": nq_step,
                   "parallelism": f"label-range sharded over {world} GPUs, one process per GPU; partial top-k rows exchanged with "
                                  f"ncclAllGather inside libwsann_cuda.so (NCCL {capi.nccl_version() if world > 1 else 'unused'}) "
                                  f"and merged on the device",
                   "l2_policy": "working set (vectors + adjacency) exceeds the 126 MB L2; no flush"},
        "e2e": {"value": round(nq_step / (e2e_ms / args.steps / 1000.0), 1), "unit": "queries/s",
                "h2d_bytes_per_step": int(nq_step * (cfg["d"] * 4 + 8)), "d2h_bytes_per_step": int(nq_step * K * 8),
                "api": "label_shard.LabelShardedTree.batch_search (host numpy in / out on every rank)"},
        "gpu_launches": int(launches), "clocks": clk,
        "device_ms_per_step": round(dev_ms, 4), "local_search_ms_per_step": round(search_ms, 4),
        "exchange_merge_ms_per_step": round(max(0.0, dev_ms - search_ms), 4),
        "exchange_bytes_per_step_per_rank": int(nq_step * K * 8),
        "hbm_gbs_peak": peaks.get("hbm_gbs"),
        "per_fraction": per_fraction,
    }
    emit(line)


class GroupRunner:
    """The public batch_search call with host buffers; behind it either one GPU or a ws_group of several."""

    def __init__(self, tree, pre, nq: int, cfg: dict, eng):
        self.tree, self.pre, self.nq, self.cfg, self.eng = tree, pre, nq, cfg, eng
        self.mean_window, self.windows, self.queries = {}, {}, None
        self.prefilter_ops = ("prefilter_tc",)  # stands for "as the engine routes it" (default options)
        self.last = None

    def upload(self, queries, windows_by_power, sorted_labels):
        self.queries = queries
        for p, w in windows_by_power.items():
            self.windows[p] = w
            self.mean_window[p] = 1e9  # route name only: the engine samples the windows itself

    def launch_dev(self, power, op):
        method, beam, mult = op
        qp = self.eng.QueryParams(K, max(1, beam), 1.35, 10_000_000, 10_000, max(1, mult), 10000, None, False)
        if method in PREFILTER_OPS:
            self.last = self.pre.batch_search(self.queries, self.windows[power], self.nq, qp)
        elif method == "super":
            self.last = self.tree.batch_search(self.queries, self.windows[power], self.nq, qp)
        else:
            self.last = self.tree.batch_search(self.queries, self.windows[power], self.nq, method, qp)

    def sync(self):
        pass

    def fetch(self):
        return self.last[0]

    def time_dev(self, power, op, reps=3):
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            self.launch_dev(power, op)
            best = min(best, (time.perf_counter() - t0) * 1000.0)
        return best


def run_group(args):
    """ONE process, ONE batch_search call, N GPUs (ws_group inside libwsann_cuda.so): the drop-in caller's view of the
    8-GPU box.  Timed end to end through the pybind classes with pageable numpy buffers, next to the same calls on one
    GPU in the same process (strong scaling of the in-engine fan-out)."""
    from rangefilteredann_b200 import capi, load_engine
    eng = load_engine()
    ndev = eng.device_count()
    if ndev == 0:
        raise SystemExit("bench.py: no CUDA device visible — this engine has no CPU fallback")
    G = args.group_gpus or ndev
    cfg = CONFIGS[args.config]
    data, queries, labels, windows = make_inputs(cfg)
    nq = len(queries)
    cdir = cache_dir(args.config)
    os.makedirs(cdir, exist_ok=True)
    validate_cache(cdir, data, labels)
    label = args.group_shard == "label"
    # the same calls on ONE GPU in the same process, for the speed-up — unless the whole data set's tree is more than
    # one GPU should be asked to build here (label shards exist for exactly that case)
    compare = not (label and cfg["n"] > 2_000_000)
    os.environ["WSANN_DEVICES"] = "0"
    os.environ.pop("WSANN_SHARD_MODE", None)
    t0 = time.time()
    tree1, pre1 = make_indices(eng, cfg, cdir, data, labels) if compare else (None, None)
    log(f"1-GPU index ready in {time.time() - t0:.1f}s" if compare else "no 1-GPU comparison at this size")
    os.environ["WSANN_DEVICES"] = args.group_devices or ",".join(str(i) for i in range(G))
    G = len(os.environ["WSANN_DEVICES"].split(","))
    if label:
        os.environ["WSANN_SHARD_MODE"] = "label"
    t0 = time.time()
    treeG, preG = make_indices(eng, cfg, cdir, data, labels)
    t_group_setup = time.time() - t0
    log(f"{G}-GPU group ({'label shards' if label else 'replicas'}) ready in {t_group_setup:.1f}s")
    gts = ground_truth_torch(data, queries, labels, windows, "cuda:0", cfg["metric"])
    sorted_labels = np.sort(labels)
    rG = GroupRunner(treeG, preG, nq, cfg, eng)
    rG.upload(queries, windows, sorted_labels)
    r1 = rG
    if compare:
        r1 = GroupRunner(tree1, pre1, nq, cfg, eng)
        r1.upload(queries, windows, sorted_labels)
    table = choose_operating_points(rG, gts, 0, cfg)
    step_methods = PREFILTER_OPS + tree_methods(cfg)
    ops = {p: min((table[p][m] for m in step_methods if m in table[p]), key=lambda v: v["ms"])["op"] for p in POWERS}

    def step(r):
        for p in POWERS:
            r.launch_dev(p, ops[p])

    res = {}
    clk = None
    for name, r in (("one_gpu", r1), ("group", rG)):
        for _ in range(args.warmup):
            step(r)
        clocks = ClockSampler(0)
        clocks.start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(r)
        res[name] = (time.perf_counter() - t0) * 1000.0 / args.steps
        c = clocks.stop()
        clk = c if name == "group" else clk
    per_fraction = {}
    identical = 0
    for p in POWERS:
        t1 = r1.time_dev(p, ops[p])
        a = r1.last
        tG = rG.time_dev(p, ops[p])
        b = rG.last
        same = bool(np.array_equal(a[0], b[0]))
        identical += int(same)
        per_fraction[frac_name(p)] = {"best": ops[p][0], "beam": ops[p][1], "final_multiply": ops[p][2],
                                      "recall": round(recall_at_k(b[0], gts[p]), 4), "recall_one_gpu": round(recall_at_k(a[0], gts[p]), 4),
                                      "ms_one_gpu": round(t1, 4), "ms_group": round(tG, 4), "speedup": round(t1 / tG, 3),
                                      "rows_identical_to_one_gpu": same}
    g = capi.Group.borrow(treeG)
    info = g.info() if g is not None else {}
    nq_step = nq * len(POWERS)
    launches = sum(g.member(i).launches() for i in range(g.size())) if g is not None else 0
    speedup = res["one_gpu"] / res["group"]
    line = {
        "metric": "QPS at recall@10>=0.95, one pass over filter fractions " + fractions_label(),
        "value": round(nq_step / (res["group"] / 1000.0), 1), "unit": "queries/s", "n_gpus": G, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(res["group"], 4), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(cfg), "name": args.config, "mode": "group/" + args.group_shard, "queries_per_step": nq_step,
                   "parallelism": f"one process, one batch_search call, {G} GPUs: " +
                                  ("label shards, partial rows gathered by peer loads over NVLink inside the merge kernel" if label
                                   else "replicas (arena cloned device to device), the batch cut into one slice per GPU, one host thread per GPU"),
                   "timing": "host wall clock around the synchronous public call (a group has no device-resident entry point), "
                             "pageable numpy in / out: value and e2e are the same measurement"},
        "e2e": {"value": round(nq_step / (res["group"] / 1000.0), 1), "unit": "queries/s",
                "h2d_bytes_per_step": int(nq_step * (cfg["d"] * 4 + 8)) * (G if label else 1), "d2h_bytes_per_step": int(nq_step * K * 8),
                "api": "pybind batch_search of the classes run_our_method.py calls, WSANN_DEVICES=" + os.environ["WSANN_DEVICES"]},
        "one_gpu": {"value": round(nq_step / (res["one_gpu"] / 1000.0), 1), "ms_per_step": round(res["one_gpu"], 4)} if compare else None,
        "speedup_vs_one_gpu": round(speedup, 3) if compare else None, "efficiency": round(speedup / G, 3) if compare else None,
        "fractions_with_rows_identical_to_one_gpu": f"{identical}/{len(POWERS)}" if compare else None,
        "group_setup_s": round(t_group_setup, 1), "group_info": info,
        "gpu_launches": int(launches), "clocks": clk, "per_fraction": per_fraction,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# the reference arm / cpu baseline
# ------------------------------------------------------------------------------------------
def ref_indices(ref, cfg, cdir, data, labels):
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the reference prints one line per loaded graph
    try:
        tree, pre = make_indices(ref, cfg, cdir, data, labels)
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
    return tree, pre


def ref_time(ref, tree, pre, queries, w, op, nq, want_dists=False):
    method, beam, mult = op
    qp = ref.QueryParams(K, max(beam, 1), 1.35, 10_000_000, 10_000, max(mult, 1), 10000, None, False)
    t0 = time.perf_counter()
    if method == "prefilter":
        ids, dd = pre.batch_search(queries[:nq], w[:nq], nq, qp)
    elif method == "super":
        ids, dd = tree.batch_search(queries[:nq], w[:nq], nq, qp)
    else:
        ids, dd = tree.batch_search(queries[:nq], w[:nq], nq, method, qp)
    dt = time.perf_counter() - t0
    return (dt, ids, dd) if want_dists else (dt, ids)


def cpu_baseline(args, cfg, data, queries, labels, windows, gts, table, ops, eng, tree_e, pre_e):
    """The reference's own implementation (oracle/_ref), all host threads, timed exactly as
    run_our_method.py does (time around one batch_search incl. argument conversion), on a
    bounded sample of each fraction's batch, at each method's operating point; per fraction
    the fastest method with recall >= 0.95 counts.

    The reference's rows are kept and compared with this engine's rows for the SAME queries, windows, graph
    files and (method, beam, final_multiply) — SURVEY.md App. G T1 / T5 at the BASELINE size, inside the bench:
      prefilter              rows equal up to distance ties within 1e-5 relative (BASELINE.json north_star)
      fenwick / opt. postf.  |recall_engine - recall_reference| <= 0.005 at equal beam
    Returns (cpu_baseline, parity)."""
    ref = load_ref()
    if ref is None:
        return {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}, None
    assert not hasattr(ref, "__engine__"), "the reference module resolved to this engine"
    cores = int(os.environ.get("PARLAY_NUM_THREADS", os.cpu_count()))
    tree, pre = ref_indices(ref, cfg, cache_dir(args.config), data, labels)
    total_q, total_t, detail = 0, 0.0, {}
    sorted_labels = np.sort(labels)
    par = {"prefilter_rows_compared": 0, "prefilter_rows_equal_up_to_1e-5_ties": 0, "prefilter_rows_bit_identical_ids": 0,
           "recall_gate": 0.005, "recall_points_compared": 0, "recall_points_within_gate": 0, "max_abs_recall_diff": 0.0,
           "per_fraction": {}}
    ref_time(ref, tree, pre, queries, windows[POWERS[len(POWERS) // 2]], ("prefilter", 0, 0), 64)  # warm the pool
    for p in POWERS:
        best = None
        pf = {}
        for method, v in table[p].items():
            if method in ("prefilter_tc", "prefilter_direct", "prefilter_auto"):  # the reference has one prefilter implementation (timed as "prefilter")
                continue
            ns = min(args.cpu_sample, len(windows[p]))
            t_probe, _ = ref_time(ref, tree, pre, queries, windows[p], v["op"], min(64, ns))
            per_q = t_probe / min(64, ns)
            ns = int(max(64, min(ns, args.cpu_budget_s / len(POWERS) / 3 / max(per_q, 1e-7))))
            t, ids, rd = ref_time(ref, tree, pre, queries, windows[p], v["op"], ns, want_dists=True)
            r = recall_at_k(ids, gts[p][:ns])
            if r >= RECALL_TARGET - 0.02 and (best is None or t / ns < best[0]):  # sample recall is noisier
                best = (t / ns, method, ns, r)
            # ---- parity on the same sample, through the same pybind classes with default options
            q_s, w_s = np.ascontiguousarray(queries[:ns]), np.ascontiguousarray(windows[p][:ns])
            _, beam, mult = v["op"]
            qp = eng.QueryParams(K, max(beam, 1), 1.35, 10_000_000, 10_000, max(mult, 1), 10000, None, False)
            if method == "prefilter":
                eids, ed = pre_e.batch_search(q_s, w_s, ns, qp)
                # windows holding fewer than k points: the reference copies k entries of a shorter frontier
                # (prefiltering.h:139-142, undefined rows) — only rows with at least k in-window points are comparable
                full = (np.searchsorted(sorted_labels, w_s[:, 1]) - np.searchsorted(sorted_labels, w_s[:, 0])) > K
                ok = rows_equal_up_to_ties(eids[full], ed[full], ids[full], rd[full], scale=1.0 if cfg.get("angular") else 0.0)
                same = int((eids[full] == ids[full]).all(1).sum())
                par["prefilter_rows_compared"] += int(full.sum())
                par["prefilter_rows_equal_up_to_1e-5_ties"] += int(ok.sum())
                par["prefilter_rows_bit_identical_ids"] += same
                pf["prefilter"] = {"rows": int(full.sum()), "rows_with_fewer_than_k_points_skipped": int((~full).sum()),
                                   "equal_up_to_ties": int(ok.sum()), "identical_ids": same}
            else:
                eids, _ = tree_e.batch_search(q_s, w_s, ns, qp) if method == "super" else tree_e.batch_search(q_s, w_s, ns, method, qp)
                r_e = recall_at_k(eids, gts[p][:ns])
                diff = abs(r_e - r)
                par["recall_points_compared"] += 1
                par["recall_points_within_gate"] += int(diff <= par["recall_gate"])
                par["max_abs_recall_diff"] = max(par["max_abs_recall_diff"], diff)
                pf[method] = {"beam": beam, "final_multiply": mult, "queries": ns, "recall_engine": round(r_e, 4),
                              "recall_reference": round(r, 4), "rows_identical_ids": int((eids == ids).all(1).sum())}
        par["per_fraction"][frac_name(p)] = pf
        if best is None:
            continue
        detail[frac_name(p)] = {"method": best[1], "qps": round(1.0 / best[0]), "sample": best[2], "recall": round(best[3], 4)}
        total_q += 1
        total_t += best[0]
    value = total_q / total_t if total_t > 0 else None  # queries/s for one query of every fraction
    par["max_abs_recall_diff"] = round(par["max_abs_recall_diff"], 4)
    par["ok"] = bool(par["prefilter_rows_compared"] == par["prefilter_rows_equal_up_to_1e-5_ties"] and
                     par["recall_points_compared"] == par["recall_points_within_gate"])
    par["against"] = "oracle/_ref (the unmodified reference) on the same queries, windows and graph files"
    if not par["ok"]:
        log("PARITY FAILURE against the reference: " + json.dumps({k2: v2 for k2, v2 in par.items() if k2 != "per_fraction"}))
    cpu = {"value": round(value, 1) if value else None, "unit": "queries/s", "cores": cores, "kind": "reference",
           "march": REF_MARCH,
           "sample": f"up to {args.cpu_sample} queries per fraction (time-bounded), same windows/graphs/operating points; "
                     f"equal weight per fraction as in the GPU step",
           "per_fraction": detail}
    return cpu, par


def run_reference(args, rank, world):
    if rank != 0:
        return
    from rangefilteredann_b200 import synth
    cfg = CONFIGS[args.config]
    ref = load_ref()
    if ref is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref not built in this snapshot"})
        return
    data, queries, labels, windows = make_inputs(cfg)
    validate_cache(cache_dir(args.config), data, labels)
    cdir = ensure_graphs(args.config, cfg, data, labels, 0)
    tree, pre = ref_indices(ref, cfg, cdir, data, labels)
    ns = args.ref_sample
    gts = {p: synth.ground_truth(data, queries[:ns], labels, windows[p][:ns], angular=cfg["metric"] == "mips") for p in POWERS}
    # operating points: smallest beam reaching the recall target per method (CPU sweep on the sample)
    ops, est_qps = {}, {}
    for p in POWERS:
        cands = [("prefilter", 0, 0)]
        for method in tree_methods(cfg):
            for beam in cfg.get("beams", BEAMS):
                _, ids = ref_time(ref, tree, pre, queries, windows[p], (method, beam, 1), ns)
                if recall_at_k(ids, gts[p]) >= RECALL_TARGET:
                    cands.append((method, beam, 1))
                    break
        timed = [(ref_time(ref, tree, pre, queries, windows[p], op, ns)[0], op) for op in cands]
        ops[p] = min(timed)[1]
        est_qps[p] = ns / min(timed)[0]
        log(f"reference {frac_name(p)}: " + ", ".join(f"{op[0]} b{op[1]} {ns / t:.0f} qps" for t, op in timed))

    # per-fraction batch sizes: large enough for parlay's fork-join to amortise on the fast
    # fractions, bounded (~0.25 s) on the slow ones
    ns_p = {p: int(min(len(windows[p]), max(ns, 0.25 * est_qps[p]))) for p in POWERS}

    def step():
        per_q = 0.0
        for p in POWERS:
            t, _ = ref_time(ref, tree, pre, queries, windows[p], ops[p], ns_p[p])
            per_q += t / ns_p[p]
        return per_q

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    per_q_sum = 0.0
    for _ in range(args.steps):
        per_q_sum += step()
    dt = time.perf_counter() - t0
    # same definition as the engine arm: equal number of queries from every fraction
    value = len(POWERS) / (per_q_sum / args.steps)
    cores = int(os.environ.get("PARLAY_NUM_THREADS", os.cpu_count()))
    line = {"impl": "reference", "metric": "QPS at recall@10>=0.95, one pass over filter fractions " + fractions_label(),
            "value": round(value, 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dt / args.steps * 1000.0, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(cfg), "name": args.config,
                       "sampling": f"each step = one batch per fraction x 17 fractions, batch = 0.25 s of work bounded to "
                                   f"[{ns}, {cfg['nq']}] queries; value = 17 / sum of per-query times"},
            "cpu_baseline": {"value": round(value, 1), "unit": "queries/s", "cores": cores, "kind": "reference",
                             "march": REF_MARCH,
                             "sample": "per fraction: " + ", ".join(f"{frac_name(p)}:{ns_p[p]}" for p in POWERS)},
            "e2e": {"value": round(value, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "per_fraction": {frac_name(p): {"method": ops[p][0], "beam": ops[p][1], "qps_estimate": round(est_qps[p])}
                             for p in POWERS}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default=os.environ.get("WSANN_BENCH_CONFIG", "auto"))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-sample", type=int, default=5000)
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--ref-sample", type=int, default=200)
    ap.add_argument("--results-csv", default=None, help="also append the operating-point sweep to this file in the "
                    "reference driver's results format (filter_width,method,recall,average_time,qps,threads)")
    ap.add_argument("--ops-file", default=None, help="save / reuse the swept operating points (profiling runs)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every rank answers its own batch (default); strong = ONE batch cut into N slices")
    ap.add_argument("--mode", default="replica", choices=["replica", "label_shard", "group"],
                    help="replica: index replicated, queries sharded over ranks (default); label_shard: every rank owns a "
                         "contiguous label range, in-library ncclAllGather + merge (one process per GPU); group: ONE process "
                         "drives --group-gpus GPUs through the public batch_search call (in-engine fan-out)")
    ap.add_argument("--group-gpus", type=int, default=0, help="--mode group: number of GPUs behind one call (0 = all)")
    ap.add_argument("--group-shard", default="replicate", choices=["replicate", "label"])
    ap.add_argument("--group-devices", default=None, help="--mode group: explicit device list, e.g. 0,0 (two members on one GPU)")
    ap.add_argument("--rows", type=int, default=0, dest="n",
                    help="scale the configuration's row count (stated in the result line); not `--n`: torchrun's own "
                         "parser claims every abbreviation of its --nnodes / --nproc-per-node in front of the script")
    ap.add_argument("--powers", default=None, help="comma-separated subset of the fractions, e.g. -12,-8,-4,0")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.config == "auto":
        args.config = "c2"  # BASELINE.json configs[1], the configuration the metric is quoted on
    global POWERS
    cfg = CONFIGS[args.config]
    if args.n and args.n != cfg["n"]:
        cfg["n"] = args.n
        cfg["scaled"] = True
    if cfg.get("adversarial"):
        POWERS = ["adv", "blowup-4", "blowup-7", "blowup-10"]
    if args.powers:
        POWERS = [(int(x) if x.lstrip("-").isdigit() else x) for x in args.powers.split(",")]
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        if args.mode == "label_shard":
            run_label_shard(args, rank, world, local_rank)
        elif args.mode == "group":
            run_group(args)
        else:
            run_engine(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
