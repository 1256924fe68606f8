#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) as markdown: one metric table per captured launch, top stall reasons,
pipe mix, and the hottest SASS instructions by stall samples.

    python tools/ncu_summary.py gpurun_out/s3/prof_sup_group_jit.ncu-rep "title" > profiles/r02_s3_ncu_hq_group_jit_supremacy30.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput, % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe, % of peak (active cycles)"),
    ("sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe, % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy, %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy, %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate, %"),
]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    print(f"# {title}\n")
    print(f"Source: `{rep}` (`ncu --set full --clock-control none`). Per-launch times under ncu are cold-cache and serialised: "
          "read shares and ratios, not absolutes.\n")
    for li, r in enumerate(raw[2:]):
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"## launch {li}: `{name}`\n")
        print("| metric | value |\n|---|---|")
        for key, label in METRICS:
            if key in hdr and r[hdr.index(key)]:
                print(f"| {label} | {r[hdr.index(key)]} {units[hdr.index(key)]} |")
        tot = float(r[hdr.index("smsp__inst_executed.sum")].replace(",", "")) if "smsp__inst_executed.sum" in hdr else 0
        st = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
        st.sort(reverse=True)
        print("\nStall reasons (warps stalled per issue-active cycle): " +
              ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st[:7]))
        pipes = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                 if h.startswith("sm__inst_executed_pipe_") and h.endswith(".sum") and r[i]]
        pipes.sort(reverse=True)
        if tot:
            print("\nPipe mix (share of executed warp instructions): " +
                  ", ".join(f"{h.replace('sm__inst_executed_pipe_', '').replace('.sum', '')} {100 * v / tot:.1f} %" for v, h in pipes[:8]))
        print()
    src = page(rep, "source", ("--print-source", "sass"))
    if len(src) > 2:
        h = src[1]
        if "Source" in h and "Warp Stall Sampling (All Samples)" in h:
            isrc, ist = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
            data = []
            for r in src[2:]:
                if r and r[0] == "Kernel Name":
                    break
                try:
                    data.append((int(r[ist]), r[isrc]))
                except (ValueError, IndexError):
                    pass
            tot = sum(d[0] for d in data) or 1
            print("## hottest SASS instructions of launch 0 (share of warp stall samples; a sample on instruction X means a warp was waiting to issue X)\n")
            print("| share | instruction (and the one before it) |\n|---|---|")
            order = sorted(range(len(data)), key=lambda i: -data[i][0])[:8]
            for i in order:
                prev = data[i - 1][1].strip() if i else ""
                print(f"| {100 * data[i][0] / tot:.1f} % | `{data[i][1].strip()[:70]}`  (after `{prev[:50]}`) |")


if __name__ == "__main__":
    main()
