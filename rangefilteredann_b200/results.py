"""Results bookkeeping of the reference's experiment driver, so that runs of this engine produce the
files the reference's own experiments/plot.py reads.  Host-side Python only; nothing
here is on the query path.

Mirrors experiments/run_our_method.py:
  compute_recall   :174-180   mean |GT ∩ top-k| / |GT| per query
  sweep_finished   :183-207   early exit of a beam sweep (the driver's `should_break` rule)
  method names     :255,295,367,397,433,525  ("prefiltering", "postfiltering_<alpha>_<beam>_<mult>",
                   "vamana-tree_<alpha>_<split>_<beam>", "optimized-postfiltering_<alpha>_<split>_<beam>_<mult>",
                   "smart-combined_…", "three-split_…", "super-postfiltering_<split>_<shift>_<alpha>_<beam>_<mult>")
  save_results     :538-567   CSV "filter_width,method,recall,average_time,qps,threads" (+ build_time,
                   branching_factor, memory columns without header, exactly as the reference writes them)
"""
from __future__ import annotations

import os

HEADER = "filter_width,method,recall,average_time,qps,threads\n"


def compute_recall(gt_neighbors, results, top_k: int) -> float:
    """run_our_method.py:174-180 (argument order as there: ground truth first)."""
    recall = 0.0
    for i in range(len(gt_neighbors)):
        gt = set(int(x) for x in gt_neighbors[i])
        res = set(int(x) for x in results[i][:top_k])
        recall += len(gt.intersection(res)) / len(gt)
    return recall / len(gt_neighbors)


def sweep_finished(rows) -> bool:
    """When the reference driver stops widening the beam for one (filter width, method) sweep
    (rule of run_our_method.py:183-207).  `rows` are the sweep's result tuples so far,
    (filter_width, method_name, recall, seconds); method names end in "_<final_beam_multiply>"."""
    if not rows:
        return False
    _, name, recall, seconds = rows[-1][:4]
    if recall > 0.999:                      # saturated
        return True
    if len(rows) < 2:
        return False
    stalled = recall <= rows[-2][2]
    if stalled and not name.endswith("_1"):  # with final_beam_multiply == 1 a stall is not conclusive
        return True
    brute_force = [r[3] for r in rows if r[1] == "prefiltering"]
    return bool(brute_force) and seconds > brute_force[-1]


def filter_width_name(power: int) -> str:
    return f"2pow{power}"                                    # run_our_method.py:29


def method_name(method: str, beam: int = 0, mult: int = 1, alpha: float = 1.0, split=2, shift=0.5) -> str:
    if method in ("prefilter", "prefiltering", "prefilter_tc", "prefilter_direct"):
        return "prefiltering"                                # :255
    if method in ("postfilter", "postfiltering", "flat"):
        return f"postfiltering_{alpha}_{beam}_{mult}"        # :295
    if method in ("fenwick", "vamana-tree"):
        return f"vamana-tree_{alpha:.3f}_{split}_{beam}"     # :367
    if method in ("optimized_postfilter", "optimized-postfiltering"):
        return f"optimized-postfiltering_{alpha:.3f}_{split}_{beam}_{mult}"  # :397
    if method in ("smart_combined", "smart-combined"):
        return f"smart-combined_{alpha:.3f}_{split}_{beam}_{mult}"           # :433
    if method in ("three_split", "three-split"):
        return f"three-split_{alpha:.3f}_{split}_{beam}_{mult}"
    if method in ("super", "super-postfiltering"):
        return f"super-postfiltering_{split}_{shift}_{alpha}_{beam}_{mult}"  # :525
    raise ValueError("unknown method " + method)


def save_results(all_results, output_file: str, num_queries: int, num_threads) -> None:
    """run_our_method.py:538-567.  Tuples are (filter_width, method, recall, total_time[, build_time
    [, branching_factor[, memory]]]); `num_threads` is what the reference prints in its `threads` column
    (here: the device, e.g. "B200x1")."""
    os.makedirs(os.path.dirname(os.path.abspath(output_file)), exist_ok=True)
    if not os.path.exists(output_file):
        with open(output_file, "a") as f:
            f.write(HEADER)
    with open(output_file, "a") as f:
        for tup in all_results:
            tup = tuple(tup) + ("",) * (7 - len(tup))
            filter_width, name, recall, total_time, build_time, branching_factor, memory = tup[:7]
            f.write(f"{filter_width},{name},{recall},{total_time / num_queries},{num_queries / total_time},{num_threads},"
                    f"{build_time},{branching_factor},{memory}\n")
