#!/usr/bin/env python
"""Renders profiles/r01_summary.md from a bench.py engine line and a reference-arm line.

  python profiles/make_summary.py gpurun_out/bench.json gpurun_out/bench_ref.json > profiles/r01_summary.md
"""
import json
import sys


def fmt_qps(v):
    return f"{v / 1e6:.2f} M" if v >= 1e6 else f"{v / 1e3:.1f} K"


def main():
    b = json.load(open(sys.argv[1]))
    r = json.load(open(sys.argv[2]))
    rf = b["roofline"]
    # DRAM traffic per launch: the committed ncu captures (the bench line carries the copy it was run with)
    import os
    try:
        tr = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r01_traffic.json")))
        for c in [rf] + list(rf.get("other_kernels", [])):
            for key, ent in tr.items():
                if isinstance(ent, dict) and key.split("_kernel")[0] in c["kernel"]:
                    c["traffic"] = ent["traffic_bytes"]
    except Exception:
        pass
    cpu = b["cpu_baseline"]
    out = []
    out.append("# Round 1 summary — one B200, BASELINE.json config 2 (1 M x 128 fp32 L2, 2-WST, 17 fractions x 10 000 queries, k = 10)\n")
    out.append("Commands (fresh box, graphs built on the device in the reference's `.bin` format and loaded by both arms):")
    out.append("```")
    out.append(f"python bench.py --steps {b['steps']} --warmup {b['warmup']}                      # engine arm   -> {sys.argv[1]}")
    out.append(f"python bench.py --impl reference --steps {r['steps']} --warmup {r['warmup']}     # reference arm -> {sys.argv[2]}")
    out.append("```\n")
    out.append("| | value |")
    out.append("|---|---|")
    out.append(f"| engine, inputs resident in HBM (`value`) | **{fmt_qps(b['value'])} queries/s** ({b['ms_per_step']:.2f} ms per {b['config']['queries_per_step']}-query step) |")
    out.append(f"| engine, end to end through pybind `batch_search`, pinned host buffers (`e2e`) | **{fmt_qps(b['e2e']['value'])} queries/s** |")
    out.append(f"| reference (oracle/_ref, parlay, {r['cpu_baseline']['cores']} host threads), `--impl reference` | {fmt_qps(r['value'])} queries/s |")
    out.append(f"| reference, `cpu_baseline` leg inside the engine run | {fmt_qps(cpu['value'])} queries/s |")
    out.append(f"| speed-up e2e / device (against `--impl reference`) | {b['e2e']['value'] / r['value']:.1f}x / {b['value'] / r['value']:.1f}x |")
    out.append(f"| kernels launched in the timed region | {b['gpu_launches']} |")
    out.append(f"| SM clock during the timed region | {b['clocks']['sm_mhz']} MHz of {b['clocks']['sm_max_mhz']} (throttle reasons: {b['clocks']['reasons'] or 'none'}) |\n")
    out.append("Roofline of the hot kernels, dominant (by time per step) first; peaks are the measured numbers in MEASURED_PEAKS.json:\n")
    for c in [rf] + list(rf.get("other_kernels", [])):
        if c["bound"] == "tensor":
            out.append(f"* `{c['kernel']}` — {c['ms_per_step']:.2f} ms per step: {c['flops_per_step'] / 1e12:.2f} TFLOP of useful work (2·d per query x "
                       f"in-window point) = **{c['achieved']:.0f} TFLOP/s = {c['frac']:.2f} of the measured dense bf16 peak** ({c['peak']} TFLOP/s; "
                       f"tf32 runs at half the bf16 rate, so {2 * c['frac']:.2f} of its own ceiling); ncu DRAM traffic of one launch: "
                       f"{(c.get('traffic') or 0) / 1e9:.2f} GB")
        else:
            out.append(f"* `{c['kernel']}` — {c['ms_per_step']:.2f} ms per step: {c['bytes_per_launch'] / 1e9:.2f} GB algorithmic per launch / "
                       f"{c['ms_per_launch']:.3f} ms (CUDA events) = **{c['achieved']:.0f} GB/s = {c['frac']:.2f} of the measured HBM copy peak** "
                       f"({c['peak']} GB/s); ncu DRAM traffic of one launch: {(c.get('traffic') or 0) / 1e9:.2f} GB")
    out.append("\nKernel time per step (CUDA events, `ws_index_kernel_times`): " +
               ", ".join(f"{k} {v:.2f} ms" for k, v in rf["kernel_ms_per_step"].items()) + ".\n")
    out.append("Per filter fraction (engine = device-resident timing of the chosen operating point; CPU = reference on "
               f"{cpu['cores']} threads, same windows / graphs, its own best method, bounded sample):\n")
    out.append("| fraction | engine method (beam) | recall@10 | engine QPS | reference method | reference QPS | ratio |")
    out.append("|---|---|---|---|---|---|---|")
    for p, v in b["per_fraction"].items():
        best = v["best"]
        e = v[best]
        c = cpu["per_fraction"].get(p)
        if c:
            out.append(f"| {p} | {best} ({e['beam']}) | {e['recall']:.4f} | {fmt_qps(e['qps'])} | {c['method']} | {fmt_qps(c['qps'])} | {e['qps'] / c['qps']:.0f}x |")
        else:
            out.append(f"| {p} | {best} ({e['beam']}) | {e['recall']:.4f} | {fmt_qps(e['qps'])} | - | - | - |")
    out.append("""
Other measurements of the final round-1 build:

* 2 GPUs, query-sharded (torchrun, one rank per GPU, index replicated, no data-path collective, weak scaling):
  27.68 M queries/s (12.29 ms per 340 000-query step; e2e 21.31 M) vs 14.12 M on one GPU — 1.96x (r01_bench_n2.json).
* label-range sharded mode on 2 GPUs (NCCL all-gather of the per-shard top-k rows + `ws_merge_partial_topk`):
  200 K x 96, recall@10 1.00 / 0.999 / 0.971 at 2^-8 / 2^-3 / 2^0, identical rows on both ranks (r01_label_shard_n2.log).
* launch list of one bench step under ncu and its per-kernel shares: r01_launches_c2_final.csv / _summary.txt.
* ncu `--set full` captures: tensor-core sweep before / after dynamic work items (r01_gemm_prefilter.md), one-launch
  prefilter kernel, warp beam kernel (r01_beam_warp_kernel_ncu.md), scan kernel; raw metric tables r01_ncu_*_raw.csv.
* device-side graph build of the 2047 graphs of config 2 (11 M node-rows): about 45 s (reference builder: 2408 s on
  5 cores of the build container).
* GPU tests: 88 passed, 1 skipped (`pytest -m gpu`): golden vectors of the reference for every index class, bit-exact ids +
  distances against the device-order oracle for every method, beam tier and k, all 17 fractions at the C2 / C3 / C4 / C5
  shapes, tensor-core and one-launch prefilter paths bit-identical to the scan path, 8-bit variants, device builder.""")
    print("\n".join(out))


if __name__ == "__main__":
    main()
