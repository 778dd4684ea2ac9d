"""Window sets shared by the golden generator and the parity tests."""
import numpy as np

from rangefilteredann_b200 import synth

TINY = dict(n=3000, d=16, nq=64, seed=7, cutoff=500)
TINY_MIPS = dict(n=2000, d=24, nq=48, seed=8, cutoff=400)  # angular; 24 floats -> rows padded to 32


def tiny_cases(labels):
    """Yields (name, windows[nq,2] f32, query-param dict)."""
    n = len(labels)
    s = np.sort(labels.astype(np.float64))
    cases = []
    # random windows of several fractions (filter_generation_utils.py recipe)
    for power in (-8, -5, -3, -1, 0):
        w = synth.make_windows(labels, power, TINY["nq"], seed=100 + power)
        # the prefilter reads past its frontier when a window holds < k points
        # (prefiltering.h:139-142), so only compare it where that cannot happen
        cases.append((f"pow{power}", w, dict(beam=20, mult=2, max_beam=10000, prefilter=power > -8)))
    # SURVEY.md Appendix D style windows expressed in ranks -> labels (midpoints)
    def mid(r):  # a value strictly between rank r-1 and rank r
        r = int(r)
        if r <= 0:
            return s[0] - 0.01
        if r >= n:
            return s[-1] + 0.01
        return 0.5 * (s[r - 1] + s[r])
    rank_windows = [(101, 2601), (701, 1601), (11, 391), (1401, 1701), (375, 1126), (0, 3000),
                    (0, 375), (2000, 2626), (1499, 1501), (5, 20), (2990, 3000), (750, 1500)]
    w = np.array([[mid(a), mid(b)] for a, b in rank_windows], dtype=np.float32)
    cases.append(("appendixD", w, dict(beam=10, mult=1, max_beam=10000, prefilter=False)))
    # small max beam: doubling stops early and slots stay padded
    w = synth.make_windows(labels, -6, 32, seed=55)
    cases.append(("maxbeam40", w, dict(beam=10, mult=1, max_beam=40, prefilter=True)))
    # ratio fallback of optimized_postfilter (range_filter_tree.h:460-466)
    w = synth.make_windows(labels, -3, 32, seed=56)
    cases.append(("ratio", w, dict(beam=20, mult=1, max_beam=10000, ratio=1.5, prefilter=True)))
    return cases


def tiny_mips_cases(labels):
    cases = []
    for power in (-6, -3, -1):
        w = synth.make_windows(labels, power, TINY_MIPS["nq"], seed=200 + power)
        cases.append((f"mips_pow{power}", w, dict(beam=10, mult=2, max_beam=10000, prefilter=True)))
    return cases


# ---- 8-bit variants (python_bindings.cpp:74-86,233-236).  64 columns: the reference's own 8-bit point
# storage needs rows that are a multiple of 64 bytes (it crashes on d = 16).
TINY_U8 = dict(n=1200, d=64, nq=32, seed=11, cutoff=600)


def tiny_u8_dataset(signed: bool):
    """Quantised Gaussian-mixture data, queries and unique labels for the 8-bit classes."""
    data, queries, labels = synth.make_dataset(TINY_U8["n"], TINY_U8["d"], TINY_U8["nq"], TINY_U8["seed"])
    if signed:
        quant = lambda x: np.clip(np.round(x * 24.0), -128, 127).astype(np.int8)
    else:
        quant = lambda x: np.clip(np.round(x * 24.0 + 128.0), 0, 255).astype(np.uint8)
    return quant(data), quant(queries), labels


def tiny_u8_cases(labels):
    cases = []
    for power in (-5, -2, 0):
        w = synth.make_windows(labels, power, TINY_U8["nq"], seed=400 + power)
        cases.append((f"u8_pow{power}", w, dict(beam=20, mult=2, max_beam=10000)))
    return cases


# ---- super-postfilter tree with a fractional split factor and a non-default shift
# (super_optimized_postfilter_tree.h:145-170 evaluates the bucket size in float)
TINY_SUPER = dict(n=800, d=16, nq=48, seed=13, cutoff=200, split=2.5, shift=0.4)


def tiny_super_cases(labels):
    cases = []
    for power in (-6, -3, -1, 0):
        w = synth.make_windows(labels, power, TINY_SUPER["nq"], seed=600 + power)
        cases.append((f"sup_pow{power}", w, dict(beam=10, mult=2, max_beam=10000)))
    return cases


# ---- duplicate labels (timestamp-style data) through PrefilterIndex: the set of in-window
# points does not depend on how equal labels are ordered, so the reference's rows are well defined
TINY_DUP = dict(n=3000, d=32, nq=64, seed=17, levels=250)


def tiny_dup_dataset():
    data, queries, _ = synth.make_dataset(TINY_DUP["n"], TINY_DUP["d"], TINY_DUP["nq"], TINY_DUP["seed"])
    rng = np.random.default_rng(TINY_DUP["seed"] + 1)
    labels = rng.integers(0, TINY_DUP["levels"], size=TINY_DUP["n"]).astype(np.float32)   # ~12 points per value
    return data, queries, labels


def tiny_dup_windows():
    """Windows whose ends sit exactly ON label values, between them, and outside the range."""
    rng = np.random.default_rng(TINY_DUP["seed"] + 2)
    nq, L = TINY_DUP["nq"], TINY_DUP["levels"]
    lo = rng.integers(0, L - 20, size=nq).astype(np.float64)
    width = rng.integers(1, 20, size=nq)
    hi = lo + width
    kind = np.arange(nq) % 4
    lo = np.where(kind == 1, lo - 0.5, lo)          # between two values
    hi = np.where(kind == 2, hi + 0.5, hi)
    w = np.stack([lo, hi], axis=1)
    w[0] = [-5.0, 3.0]                              # starts below every label
    w[1] = [L - 4.0, L + 10.0]                      # ends above every label (r = n-1 rule: last point dropped)
    w[2] = [0.0, float(L)]                          # everything
    w[3] = [7.0, 7.0]                               # empty: lo == hi on a label value
    return w.astype(np.float32)
