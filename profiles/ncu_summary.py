#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per kernel launch the headline metrics, the warp-stall
breakdown and the SASS lines with the most stall samples.

    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [--top 12]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("====", r[h.index("Kernel Name")][:90])
        for k in KEYS:
            if k in h:
                print(f"  {k}: {r[h.index(k)]} {units[h.index(k)]}")
        st = [(float(r[i] or 0), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
              for i, k in enumerate(h) if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
        print("  stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        if not b["rows"]:
            continue
        hdr = b["rows"][0]
        ia, it, isrc, ismp = (hdr.index(x) for x in ("Instructions Executed", "Avg. Threads Executed", "Source", "# Samples"))
        stall_cols = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
        data = [(int(r[ismp] or 0), int(r[ia] or 0), float(r[it] or 0), r[isrc].strip(), k, r)
                for k, r in enumerate(b["rows"][1:]) if len(r) > ia]
        tsm, tin = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
        print("==== hot SASS (by stall samples):", b["name"][:70], f"[{tin} warp instr, {len(data)} SASS lines]")
        for d in sorted(data, reverse=True)[:top]:
            st = sorted([(int(d[5][i] or 0), hdr[i][6:]) for i in stall_cols], reverse=True)[:2]
            print(f"  {d[0] / tsm * 100:5.1f}% smp  exec {d[1]:>10d}  thr {d[2]:4.1f}  #{d[4]:<5d} {d[3][:52]:52s} {st}")


if __name__ == "__main__":
    main()
