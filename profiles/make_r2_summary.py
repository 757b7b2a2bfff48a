#!/usr/bin/env python
"""Builds profiles/r2_ncu_summary.md, r2_ncu_full_step.json and r2_dram_traffic.json from the round-2 captures
(launch list CSV + the two --set full reports).  Run here (no GPU): python profiles/make_r2_summary.py"""
import collections
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = {"gpu__time_duration.sum": "duration_ns", "launch__grid_size": "grid", "launch__block_size": "block",
        "launch__registers_per_thread": "regs", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "smsp__inst_executed.sum": "warp_inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts", "lts__t_sectors.sum": "l2_sectors",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak"}


def short(name):
    return name.split("(")[0].replace("void ", "").replace("arb::", "")


def launch_list():
    rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r2_launches_step.csv"))) if len(r) > 10 and r[0].isdigit()]
    L = collections.OrderedDict()
    for r in rows:
        L.setdefault(int(r[0]), {"name": short(r[4]), "grid": r[8], "block": r[7]})[r[12]] = float(r[14].replace(",", ""))
    fuse = [i for i in L if "fuse_kernel" in L[i]["name"]]
    lo, hi = fuse[-2] + 1, fuse[-1]
    agg = collections.OrderedDict()
    for i in range(lo, hi + 1):
        e = L[i]
        a = agg.setdefault(e["name"], {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0, "grid": e["grid"], "block": e["block"]})
        a["launches"] += 1
        a["us"] += e.get("gpu__time_duration.sum", 0) / 1e3
        a["dram_read"] += e.get("dram__bytes_read.sum", 0)
        a["dram_write"] += e.get("dram__bytes_write.sum", 0)
    return agg


def full(rep):
    txt = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": short(r[h.index("Kernel Name")])}
        for k, nm in KEYS.items():
            if k in h:
                v = float(r[h.index(k)].replace(",", "") or 0)
                u = units[h.index(k)]
                mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "ms": 1e6, "us": 1e3, "s": 1e9}.get(u, 1.0)
                d[nm] = v * mult
        st = sorted([(float(r[i] or 0), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for i, k in enumerate(h) if "issue_stalled" in k and k.endswith("per_issue_active.ratio")], reverse=True)[:5]
        d["stalls_per_issue"] = {n: round(v, 2) for v, n in st}
        out.append(d)
    return out


def main():
    agg = launch_list()
    total = sum(a["us"] for a in agg.values())
    fa, fb = full("r2_step_full_a.ncu-rep"), full("r2_step_full_b.ncu-rep")
    json.dump({"launch_list_of_one_step": agg, "step_total_us_under_ncu": total, "set_full": fa + fb,
               "source": "ncu 2025 on one B200 via gpurun; tests/gpu_profile_step.py (cfg3: 1 M notes, 4096 queries); second of two steps"},
              open(os.path.join(ROOT, "profiles", "r2_ncu_full_step.json"), "w"), indent=1)

    def grp(pred):
        return int(sum(a["dram_read"] + a["dram_write"] for k, a in agg.items() if pred(k)))
    traffic = {"dense_topk": grp(lambda k: k.startswith("dense_topk")), "bm25_topk": grp(lambda k: k.startswith("bm25_")),
               "bm25_member": grp(lambda k: "member" in k or "compact" in k or "bt_" in k or "qclassify" in k),
               "bm25_screen": grp(lambda k: "screen" in k or "recheck" in k or "h3" in k),
               "bm25_final": grp(lambda k: "final" in k or "bm25_topk_kernel" in k or "flag_merge" in k),
               "fuse": grp(lambda k: k.startswith("fuse")),
               "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the launches of one cfg3 step "
                         "(profiles/r2_launches_step.csv, the build of profiles/r2_bench_n1.json)"}
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "r2_dram_traffic.json"), "w"), indent=1)
    md = ["# Round-2 profiling evidence (B200, sm_100a)", "",
          "Captured with `gpurun` on one B200 after the profiled command had exited 0 without ncu. Per-launch times under ncu are",
          "cold-cache and serialised: compare SHARES with the live CUDA-event stage times of `profiles/r2_bench_n1.json`",
          "(one-stream pass: dense 4.76 ms, BM25 4.24 ms = 3.00 member pass + 0.92 screening + 0.29 final/fallback, fuse 0.17 ms = 9.19 ms;",
          "headline with the two sources on two streams: 8.70-9.01 ms over five runs of the final builds).", "",
          "## 1. Launch list of one cfg3 step", "",
          "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv python tests/gpu_profile_step.py`",
          "(raw: `profiles/r2_launches_step.csv`; the second of the two steps the script runs)", "",
          "| kernel | launches | grid x block | us | share | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|---|"]
    for k, a in agg.items():
        md.append(f"| `{k}` | {a['launches']} | {a['grid']} x {a['block']} | {a['us']:.1f} | {a['us'] / total * 100:.1f} % | {a['dram_read'] / 1e6:.1f} | {a['dram_write'] / 1e6:.1f} |")
    md += [f"| **step** | | | **{total:.1f}** | 100 % | {sum(a['dram_read'] for a in agg.values()) / 1e6:.0f} | {sum(a['dram_write'] for a in agg.values()) / 1e6:.0f} |", "",
           "## 2. `ncu --set full --import-source on --clock-control none` of the step's kernels", "",
           "(reports: `gpurun_out/r2_step_full_a.ncu-rep` = dense / screening / fuse, `r2_step_full_b.ncu-rep` = last member-pass round;",
           "machine-readable copy: `profiles/r2_ncu_full_step.json`)", "",
           "| kernel | duration | regs | warps active | issue active | tensor pipe | smem wavefronts (conflicts) | L2 sectors (hit) | DRAM rd / wr | top stalls per issue |",
           "|---|---|---|---|---|---|---|---|---|---|"]
    for d in fa + fb:
        md.append(f"| `{d['kernel']}` | {d['duration_ns'] / 1e6:.3f} ms | {int(d['regs'])} | {d['warps_active_pct']:.1f} % | {d['issue_active_pct']:.1f} % | "
                  f"{d.get('tensor_pipe_active_pct', 0):.1f} % | {d['smem_wavefronts'] / 1e6:.1f} M ({d['smem_bank_conflicts'] / 1e6:.1f} M) | "
                  f"{d['l2_sectors'] / 1e6:.0f} M ({d['l2_hit_pct']:.0f} %) | {d['dram_read'] / 1e6:.0f} / {d['dram_write'] / 1e6:.0f} MB | "
                  f"{', '.join(f'{k} {v}' for k, v in d['stalls_per_issue'].items())} |")
    md += ["", "## 3. History of the step this round (cfg3, one B200, live CUDA-event timing of `bench.py`)", "",
           "| change | BM25 stage | dense stage | step |", "|---|---|---|---|",
           "| end of round 1 (per-(query, part) sparse pass + fp64 plane pass) | 8.83 ms | 4.77 | 13.86 ms |",
           "| tile-major join: member pass + fp32 upper-bound screening pass with queued fp64 rechecks | 7.4 | | |",
           "| term-major tile rows, records decoded once per run, per-doc filters, set-aside pairs in their own kernel | 6.4 | | |",
           "| histogram-select compaction, allocation-based query-term index | 5.8 | | |",
           "| per-doc filters precomputed in the index, parallel tile fill | 5.72 | 4.74 | 10.65 |",
           "| member pass screens pairs in fp32 (rounded-up plane tiles, upward-rounded in-order sums); exact fp64 only above the bound | 5.41 | | |",
           "| query-term records stored decoded (decode ran with a third of the lanes) | 5.05 | | |",
           "| exact kernel templated on the query width (plane loads overlap), early duplicate exit | 4.92 | 4.87 | 9.95 |",
           "| dense and BM25 on two streams | | | 9.16 |",
           "| screening pass: warps take 8 active queries at a time instead of 32 | 4.79 | | |",
           "| one record for a sparse term repeated inside a query (13.1 M -> 8.8 M exact evaluations per batch) | 4.56 | | |",
           "| member pass at 768 threads per CTA (80 registers) | 4.32 | 4.73 | 8.93 (9.24 on one stream) |",
           "| rounds grow x4 throughout (no extra half-corpus round) | 4.24 | 4.76 | 8.81 (9.19 on one stream) |",
           "| run structure of the tile rows precomputed in the index; BM25 lists of the two-stream path in persistent buffers | 4.18-4.24 | 4.75-4.81 | 8.70-9.01 (9.12-9.25) |", "",
           "Negative results: tiles split into more units to fill whole waves (3.52 -> 3.77 ms member pass: every unit re-stages its tile);",
           "queue entries carrying doc / contribution instead of a posting index (no change: the exact kernel is bound by the plane and row sectors).",
           "cfg5 (10 M docs): BM25 stage 158 -> 69.6 ms by sizing the rounds in tiles (the 64ths plan overflowed the candidate lists in round 1,",
           "90 ms of per-query fallback), then -> 35.7 ms with launches bounded to 2048 tiles (the exact kernel's queues overflowed) and the",
           "fp32-screened member pass, -> 29.8 ms with the later member-pass changes; cfg5 step 162 -> 33.0 ms.", "",
           "Also measured slower: 6 / 8 postings per thread in the join's preparation (3.56 / 4.31 ms member pass vs 3.09), 832 / 1024 threads per",
           "member-pass CTA (3.25 / 4.24 ms; register spills), screening grabs of 2 / 4 / 16 queries (1.18 / 0.98 / 0.96 ms vs 0.91 at 8); a sampled-",
           "histogram bound for every bound-less query at 4096 queries (+0.09 ms: their low thresholds add screening work while the per-query",
           "kernel still runs for the overflowed ones) -- kept for batches <= 2048 queries where it removes the per-query kernel (-0.2 ms).", "",
           "Kernel-only timing experiments (ARB_BM25_DEBUG, wrong results, timing only): member pass with pairs skipped 0.60 ms (4096 queries) /",
           "0.56 ms (512) = the fixed cost; with the queue disabled 2.30 / 0.97 ms; complete 3.66 / 1.26 ms."]
    open(os.path.join(ROOT, "profiles", "r2_ncu_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
