"""CPU: host-side logic of bench.py that the multi-GPU runs depend on.

The graph cache is keyed by file names that only carry label VALUES (postfilter_vamana.h:126-132), so
every rank, every world size and the reference arm must see the same points and labels; only the
queries differ between ranks.  (A 2-GPU run once generated its data with nq * world queries, which
shifted the label stream: the reference arm then loaded graphs of another point set.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from rangefilteredann_b200 import synth  # noqa: E402


def test_rank_queries_do_not_touch_the_dataset():
    d0, q0, l0 = synth.make_dataset(5000, 16, 100, seed=0)
    d1, q1, l1 = synth.make_dataset(5000, 16, 100, seed=0)
    assert np.array_equal(d0, d1) and np.array_equal(l0, l1) and np.array_equal(q0, q1)
    r1 = synth.make_rank_queries(16, 100, 0, rank=1)
    r1b = synth.make_rank_queries(16, 100, 0, rank=1)
    r2 = synth.make_rank_queries(16, 100, 0, rank=2)
    assert r1.shape == q0.shape and r1.dtype == np.float32
    assert np.array_equal(r1, r1b) and not np.array_equal(r1, r2) and not np.array_equal(r1, q0)
    # same mixture as the data: every query sits within a few sigma (0.5 per component) of one of the 256 centres
    centers = np.random.default_rng(0).standard_normal((256, 16)).astype(np.float32)
    dmin = np.sqrt(((r1[:, None, :] - centers[None, :, :]) ** 2).sum(2).min(1))
    assert dmin.max() < 0.5 * np.sqrt(16) * 2.0
    ang = synth.make_rank_queries(16, 50, 0, rank=3, angular=True)
    assert np.allclose(np.linalg.norm(ang, axis=1), 1.0, atol=1e-5)


def test_cache_fingerprint(tmp_path):
    data, _, labels = synth.make_dataset(2000, 8, 4, seed=1)
    cdir = str(tmp_path / "wst") + "/"
    os.makedirs(cdir)
    open(os.path.join(cdir, "vamana_500_64_1.000000_0.0_1.0_2000.bin"), "wb").write(b"x")
    bench.validate_cache(cdir, data, labels)            # cache of unknown origin: dropped
    assert [f for f in os.listdir(cdir) if f.endswith(".bin")] == []
    open(os.path.join(cdir, "g.bin"), "wb").write(b"x")
    bench.validate_cache(cdir, data, labels)            # same dataset: kept
    assert os.path.exists(os.path.join(cdir, "g.bin"))
    _, _, other = synth.make_dataset(2000, 8, 5, seed=1)  # one more query shifts the label stream
    assert not np.array_equal(other, labels) and np.array_equal(np.sort(other), np.sort(labels))
    bench.validate_cache(cdir, data, other)             # same label values, other assignment: dropped
    assert not os.path.exists(os.path.join(cdir, "g.bin"))


def test_expected_graph_count():
    assert bench.expected_graph_count(1_000_000, 1000) == 2047   # 11 rows
    assert bench.expected_graph_count(3000, 500) == 15
