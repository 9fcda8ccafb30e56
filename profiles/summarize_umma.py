#!/usr/bin/env python
"""Summaries of the tcgen05 selector captures (ncu --set full of conv_umma_kernel launches) under profiles/.

    python profiles/summarize_umma.py gpurun_out/r4c_umma.ncu-rep r2_umma_l2 [launch index = 0]

Writes profiles/<name>_kernel.json (selected metrics of that launch) and profiles/<name>_stalls.txt (warp-stall samples per
warp role: the roles are told apart by the SASS they execute -- LDTM = epilogue, UTCHMMA = MMA issue, UTMALDG = TMA producer,
LDS/STS.128 split loop or LDG gather = operand builders).  Reads the report offline (no GPU).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_uniform_realtime.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, name = sys.argv[1], sys.argv[2]
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2 + idx]
    out = {"report": os.path.basename(rep), "launch": idx, "kernel": vals[hdr.index("Kernel Name")]}
    for i, h in enumerate(hdr):
        short = h.split("TriageCompute.")[-1]
        if short in KEYS and vals[i] != "":
            out[short] = {"value": vals[i], "unit": units[i]}
    json.dump(out, open(os.path.join(ROOT, "profiles", name + "_kernel.json"), "w"), indent=1)

    src = ncu_csv(rep, "source", ("--print-source", "sass"))
    # the source page lists every captured kernel one after the other: take the idx-th block
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(r)
    blk = blocks[idx]
    h2 = blk[0]
    ix = {h: i for i, h in enumerate(h2)}
    data = blk[1:]
    reasons = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
    # split the SASS into role regions at the role-defining instructions
    marks = {"epilogue": "LDTM", "mma issue": "UTCHMMA", "tma producer": "UTMALDG"}
    first = {k: next((n for n, r in enumerate(data) if m in r[ix["Source"]]), None) for k, m in marks.items()}
    t = out.get("gpu__time_duration.sum", {})
    lines = [f"{out['kernel']}", f"gpu__time_duration {t.get('value')} {t.get('unit')}", ""]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    lines.append(f"warp-stall samples: {tot} in total; by reason (all warps):")
    agg = {k: sum(int(r[ix[k]]) for r in data) for k in reasons}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        if v:
            lines.append(f"  {k:28s} {v:8d}  {100.0 * v / max(tot, 1):5.1f} %")
    lines.append("")
    lines.append("top instructions by samples (address order kept):")
    top = sorted(range(len(data)), key=lambda n: -int(data[n][ix["# Samples"]]))[:24]
    for n in sorted(top):
        r = data[n]
        st = {k[6:]: int(r[ix[k]]) for k in reasons if int(r[ix[k]]) > 0}
        lines.append(f"  #{n:5d} {int(r[ix['# Samples']]):7d} samples  {int(r[ix['Instructions Executed']]):10d} exec  {r[ix['Source']].strip()[:64]:64s} {st}")
    lines.append("")
    lines.append("first SASS index of the role-defining instructions: " + ", ".join(f"{k}: {v}" for k, v in first.items()))
    open(os.path.join(ROOT, "profiles", name + "_stalls.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main()
