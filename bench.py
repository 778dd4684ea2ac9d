#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config, per the driver contract.

  python bench.py --gpus N --steps K --warmup W            this engine (one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU path

Metric: QPS at recall@10 >= 0.95 over the 17 filter fractions 2^-16..2^0 (k = 10, 10 000
queries per fraction).  One "step" = one pass over all 17 fractions: for each fraction the
batch of `nq` queries is answered by the fastest (method, beam, final_multiply) operating
point that reaches recall@10 >= 0.95 — methods: prefilter (task path, the one-launch kernel
"prefilter_direct", or the tensor-core sweep "prefilter_tc" — same rows), range-filter tree ("fenwick"), optimized postfilter — chosen
in an untimed sweep, exactly the pareto rule of the
reference's plots (experiments/plot.py:14-28).  value = queries answered per second.

  value      inputs resident in HBM, CUDA-event timed on the engine's stream
  e2e        same step through the public pybind `batch_search` with pinned HOST buffers
             (H2D of queries+windows and D2H of ids+dists inside the timed region)
  roofline   dominant kernel, algorithmic bytes / CUDA-event kernel time
  cpu_baseline  the unmodified reference (oracle/_ref) on the box's host cores, bounded sample

Multi-GPU (torchrun): index replicated, every rank answers its own batch (weak scaling, no
data-path collective); time = max over ranks.
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
    "c4": dict(n=12_000_000, d=512, nq=10_000, seed=0, cutoff=1000, metric="mips", angular=True, labels="timestamp",
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
    labels = {"timestamp": "timestamp-style integer labels (duplicates)"}.get(cfg.get("labels", ""), "uniform unique labels")
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


def rows_equal_up_to_ties(ids, dists, rids, rdists, rtol=1e-5):
    """Per row: ids equal position-wise, except among entries whose distances tie within rtol (BASELINE.json:
    'prefilter top-k ids must match the reference exactly, except for distance ties within 1e-5 relative')."""
    ok = np.zeros(len(ids), dtype=bool)
    for i in range(len(ids)):
        if not np.allclose(dists[i], rdists[i], rtol=rtol, atol=1e-30):
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

    def __init__(self, tree_index, nq: int, d: int):
        from rangefilteredann_b200 import capi
        self.capi = capi
        self.tree = tree_index
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


def choose_operating_points(runner: EngineRunner, gts, rank):
    """Untimed sweep: per fraction and method, the first (smallest) beam reaching the recall
    target for each final_multiply; the fastest of those is the method's operating point."""
    table = {}
    for p in POWERS:
        per_method = {}
        # prefilter: exact
        runner.launch_dev(p, ("prefilter", 0, 0))
        r = recall_at_k(runner.fetch(), gts[p])
        ms = runner.time_dev(p, ("prefilter", 0, 0))
        per_method["prefilter"] = dict(op=("prefilter", 0, 0), recall=r, ms=ms)
        for alt in ("prefilter_direct", "prefilter_tc"):
            runner.launch_dev(p, (alt, 0, 0))
            r = recall_at_k(runner.fetch(), gts[p])
            ms = runner.time_dev(p, (alt, 0, 0))
            per_method[alt] = dict(op=(alt, 0, 0), recall=r, ms=ms)
        # what PrefilterIndex.batch_search picks by itself for this fraction's windows (the step uses THIS, not the
        # fastest of the three)
        per_method["prefilter_auto"] = dict(per_method[auto_prefilter_route(runner.mean_window[p])])
        for method in ("fenwick", "optimized_postfilter"):
            best = None
            for mult in (MULTS if method == "optimized_postfilter" else [1]):
                for beam in BEAMS:
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
            log(f"2^{p}: " + ", ".join(f"{m}: beam {v['op'][1]} x{v['op'][2]} recall {v['recall']:.4f} {v['ms']:.3f} ms"
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
    # cache is keyed by them); each rank answers its own batch of queries (weak scaling)
    data, queries, labels = synth.make_dataset(cfg["n"], cfg["d"], cfg["nq"], cfg["seed"])
    if rank > 0:
        queries = synth.make_rank_queries(cfg["d"], cfg["nq"], cfg["seed"], rank)
    from rangefilteredann_b200 import sharding
    windows = {p: synth.make_windows(labels, p, cfg["nq"], seed=1000 + 17 * rank + p) for p in POWERS}
    cdir = cache_dir(args.config)
    os.makedirs(cdir, exist_ok=True)
    if rank == 0:
        validate_cache(cdir, data, labels)
    if world > 1:  # rank 0 fills the cache (its constructor builds + saves what is missing), the others load it
        import torch.distributed as dist
        if rank == 0:
            ensure_graphs(args.config, cfg, data, labels, 0)
        dist.barrier()
    tree = eng.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, cfg["cutoff"], 2, eng.BuildParams(64, 500, 1.0, cdir))
    log(f"rank {rank}: data + index ready in {time.time() - t_setup:.1f}s")
    gts = ground_truth_torch(data, queries, labels, windows, f"cuda:{local_rank}")
    pre = eng.PrefilterIndexFloatEuclidian(data, labels)  # the class run_our_method.py:249 calls for "prefiltering"
    sorted_labels = np.sort(labels)
    runner = EngineRunner(tree, cfg["nq"], cfg["d"])
    runner.upload(queries, windows, sorted_labels)
    h = runner.h
    if args.ops_file and os.path.exists(args.ops_file):
        saved = json.load(open(args.ops_file))
        table = {int(p): {m: dict(op=tuple(v["op"]), recall=v["recall"], ms=v["ms"]) for m, v in pm.items()}
                 for p, pm in saved.items()}
    else:
        table = choose_operating_points(runner, gts, rank)
        if args.ops_file and rank == 0:
            json.dump({str(p): {m: dict(op=list(v["op"]), recall=v["recall"], ms=v["ms"]) for m, v in pm.items()}
                       for p, pm in table.items()}, open(args.ops_file, "w"))
    # per fraction: the fastest of {prefilter as the engine routes it by itself, range-filter tree, optimized postfilter}
    step_methods = ("prefilter_auto", "fenwick", "optimized_postfilter")
    ops = {p: min((table[p][m] for m in step_methods if m in table[p]), key=lambda v: v["ms"])["op"] for p in POWERS}
    nq_step = cfg["nq"] * len(POWERS)

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
            return pre.batch_search(queries, windows[p], cfg["nq"], qps_by_op.setdefault((10, 1), eng.QueryParams(K, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False)))
        qp = qps_by_op.setdefault((beam, mult), eng.QueryParams(K, beam, 1.35, 10_000_000, 10_000, mult, 10000, None, False))
        return tree.batch_search(queries, windows[p], cfg["nq"], method, qp)

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
            routes[f"2^{p}"] = "prefilter_tc" if "gemm_sweep" in kt else ("prefilter" if "decompose" in kt else "prefilter_direct")
            if routes[f"2^{p}"] != ops[p][0]:
                log(f"WARNING 2^{p}: auto routing took {routes[f'2^{p}']}, the device-timed step ran {ops[p][0]}")
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
    value = world * nq_step / (ms_per_step / 1000.0)
    e2e_value = world * nq_step / (e2e_ms / args.steps / 1000.0)

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    dpad_bytes = ((cfg["d"] * 4 + 63) // 64) * 64
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
        gemm_info = {"kernel": "ws_gemm_topk_kernel (tcgen05 kind::tf32)", "bound": "tensor",
                     "achieved": round(gemm_flops / (gemm_ms / 1000.0) / 1e12, 1), "peak": tensor_peak, "unit": "TFLOP/s",
                     "frac": round(gemm_flops / (gemm_ms / 1000.0) / 1e12 / tensor_peak, 4),
                     "peak_source": "measured dense bf16 (MEASURED_PEAKS.json bf16_tflops); the kernel runs tf32, whose "
                                    "dense rate is half of bf16",
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
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
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
        per_fraction[f"2^{p}"] = {m: {"beam": v["op"][1], "final_multiply": v["op"][2], "recall": round(v["recall"], 4),
                                      "qps": round(cfg["nq"] / (v["ms"] / 1000.0))} for m, v in table[p].items()}
        per_fraction[f"2^{p}"]["best"] = ops[p][0]
    line = {
        "metric": "QPS at recall@10>=0.95, one pass over filter fractions 2^-16..2^0",
        "value": round(value, 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(cfg),
                   "name": args.config, "queries_per_step": nq_step * world,
                   "l2_policy": "working set (vectors + adjacency, >= 0.25 GB, random gathers) exceeds the 126 MB L2; no flush",
                   "parallelism": f"query-sharded dp{world}, index replicated"},
        "e2e": {"value": round(e2e_value, 1), "unit": "queries/s",
                "h2d_bytes_per_step": int(len(POWERS) * cfg["nq"] * (cfg["d"] * 4 + 8)),
                "d2h_bytes_per_step": int(len(POWERS) * cfg["nq"] * K * 8),
                "api": "pybind PrefilterIndexFloatEuclidian.batch_search / VamanaRangeFilterTreeIndexFloatEuclidian.batch_search, "
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
# the reference arm / cpu baseline
# ------------------------------------------------------------------------------------------
def ref_indices(ref, cfg, cdir, data, labels):
    bp = ref.BuildParams(64, 500, 1.0, cdir)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the reference prints one line per loaded graph
    try:
        tree = ref.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, cfg["cutoff"], 2, bp)
        pre = ref.PrefilterIndexFloatEuclidian(data, labels)
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
    par = {"prefilter_rows_compared": 0, "prefilter_rows_equal_up_to_1e-5_ties": 0, "prefilter_rows_bit_identical_ids": 0,
           "recall_gate": 0.005, "recall_points_compared": 0, "recall_points_within_gate": 0, "max_abs_recall_diff": 0.0,
           "per_fraction": {}}
    ref_time(ref, tree, pre, queries, windows[-8], ("prefilter", 0, 0), 64)  # warm the pool
    for p in POWERS:
        best = None
        pf = {}
        for method, v in table[p].items():
            if method in ("prefilter_tc", "prefilter_direct", "prefilter_auto"):  # the reference has one prefilter implementation (timed as "prefilter")
                continue
            ns = min(args.cpu_sample, cfg["nq"])
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
                ok = rows_equal_up_to_ties(eids, ed, ids, rd)
                same = int((eids == ids).all(1).sum())
                par["prefilter_rows_compared"] += ns
                par["prefilter_rows_equal_up_to_1e-5_ties"] += int(ok.sum())
                par["prefilter_rows_bit_identical_ids"] += same
                pf["prefilter"] = {"rows": ns, "equal_up_to_ties": int(ok.sum()), "identical_ids": same}
            else:
                eids, _ = tree_e.batch_search(q_s, w_s, ns, method, qp)
                r_e = recall_at_k(eids, gts[p][:ns])
                diff = abs(r_e - r)
                par["recall_points_compared"] += 1
                par["recall_points_within_gate"] += int(diff <= par["recall_gate"])
                par["max_abs_recall_diff"] = max(par["max_abs_recall_diff"], diff)
                pf[method] = {"beam": beam, "final_multiply": mult, "queries": ns, "recall_engine": round(r_e, 4),
                              "recall_reference": round(r, 4), "rows_identical_ids": int((eids == ids).all(1).sum())}
        par["per_fraction"][f"2^{p}"] = pf
        if best is None:
            continue
        detail[f"2^{p}"] = {"method": best[1], "qps": round(1.0 / best[0]), "sample": best[2], "recall": round(best[3], 4)}
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
    data, queries, labels = synth.make_dataset(cfg["n"], cfg["d"], cfg["nq"], cfg["seed"])
    windows = {p: synth.make_windows(labels, p, cfg["nq"], seed=1000 + p) for p in POWERS}
    validate_cache(cache_dir(args.config), data, labels)
    cdir = ensure_graphs(args.config, cfg, data, labels, 0)
    tree, pre = ref_indices(ref, cfg, cdir, data, labels)
    ns = args.ref_sample
    gts = {p: synth.ground_truth(data, queries[:ns], labels, windows[p][:ns]) for p in POWERS}
    # operating points: smallest beam reaching the recall target per method (CPU sweep on the sample)
    ops, est_qps = {}, {}
    for p in POWERS:
        cands = [("prefilter", 0, 0)]
        for method in ("fenwick", "optimized_postfilter"):
            for beam in BEAMS:
                _, ids = ref_time(ref, tree, pre, queries, windows[p], (method, beam, 1), ns)
                if recall_at_k(ids, gts[p]) >= RECALL_TARGET:
                    cands.append((method, beam, 1))
                    break
        timed = [(ref_time(ref, tree, pre, queries, windows[p], op, ns)[0], op) for op in cands]
        ops[p] = min(timed)[1]
        est_qps[p] = ns / min(timed)[0]
        log(f"reference 2^{p}: " + ", ".join(f"{op[0]} b{op[1]} {ns / t:.0f} qps" for t, op in timed))

    # per-fraction batch sizes: large enough for parlay's fork-join to amortise on the fast
    # fractions, bounded (~0.25 s) on the slow ones
    ns_p = {p: int(min(cfg["nq"], max(ns, 0.25 * est_qps[p]))) for p in POWERS}

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
    line = {"impl": "reference", "metric": "QPS at recall@10>=0.95, one pass over filter fractions 2^-16..2^0",
            "value": round(value, 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dt / args.steps * 1000.0, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(cfg), "name": args.config,
                       "sampling": f"each step = one batch per fraction x 17 fractions, batch = 0.25 s of work bounded to "
                                   f"[{ns}, {cfg['nq']}] queries; value = 17 / sum of per-query times"},
            "cpu_baseline": {"value": round(value, 1), "unit": "queries/s", "cores": cores, "kind": "reference",
                             "march": REF_MARCH,
                             "sample": "per fraction: " + ", ".join(f"2^{p}:{ns_p[p]}" for p in POWERS)},
            "e2e": {"value": round(value, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "per_fraction": {f"2^{p}": {"method": ops[p][0], "beam": ops[p][1], "qps_estimate": round(est_qps[p])}
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
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.config == "auto":
        args.config = "c2"  # BASELINE.json configs[1], the configuration the metric is quoted on
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        run_engine(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
