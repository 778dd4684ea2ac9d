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


def test_results_csv_matches_reference_driver_format(tmp_path):
    """rangefilteredann_b200/results.py against the reference driver's conventions
    (experiments/run_our_method.py:174-207,538-567)."""
    from rangefilteredann_b200 import results as res
    gt = np.array([[1, 2, 3, 4], [5, 6, 7, 8]])
    out = np.array([[1, 2, 9, 9, 3], [8, 7, 6, 5, 0]])
    assert res.compute_recall(gt, out, 4) == (2 / 4 + 4 / 4) / 2
    assert res.method_name("prefilter_tc") == "prefiltering"
    assert res.method_name("fenwick", 20) == "vamana-tree_1.000_2_20"
    assert res.method_name("optimized_postfilter", 80, 2) == "optimized-postfiltering_1.000_2_80_2"
    assert res.method_name("super", 10, 4) == "super-postfiltering_2_0.5_1.0_10_4"
    assert res.filter_width_name(-8) == "2pow-8"
    # sweep_finished: recall ~ 1 stops; no improvement stops unless final multiply is 1; slower than prefiltering stops
    assert not res.sweep_finished([])
    assert res.sweep_finished([("2pow-8", "x_1", 0.9995, 1.0)])
    assert not res.sweep_finished([("2pow-8", "x_10_1", 0.9, 1.0)])
    assert res.sweep_finished([("w", "x_10_2", 0.9, 1.0), ("w", "x_20_2", 0.9, 1.0)])
    assert not res.sweep_finished([("w", "x_10_1", 0.9, 1.0), ("w", "x_20_1", 0.9, 1.0)])
    assert res.sweep_finished([("w", "prefiltering", 1.0 - 1e-3 - 1e-9, 0.5), ("w", "x_10_1", 0.8, 0.1), ("w", "x_20_1", 0.9, 0.7)])
    path = str(tmp_path / "results" / "sift_results.csv")
    res.save_results([("2pow-8", "prefiltering", 1.0, 2.0), ("2pow-8", "vamana-tree_1.000_2_10", 0.97, 0.5, 12.5, 2, 100)],
                     path, 10000, "B200x1")
    res.save_results([("2pow0", "prefiltering", 1.0, 4.0)], path, 10000, "B200x1")
    lines = open(path).read().splitlines()
    assert lines[0] == "filter_width,method,recall,average_time,qps,threads"
    assert lines[1] == "2pow-8,prefiltering,1.0,0.0002,5000.0,B200x1,,,"
    assert lines[2] == "2pow-8,vamana-tree_1.000_2_10,0.97,5e-05,20000.0,B200x1,12.5,2,100"
    assert lines[3].startswith("2pow0,prefiltering,1.0,0.0004,2500.0") and len(lines) == 4
    import pandas as pd  # the reference's plot.py reads the file with pandas
    df = pd.read_csv(path, index_col=False, usecols=range(6))
    assert list(df.columns) == ["filter_width", "method", "recall", "average_time", "qps", "threads"] and len(df) == 3
