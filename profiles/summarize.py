#!/usr/bin/env python
"""Turn one gpurun profiling call (scratch/run_prof.sh) into the tracked summaries under profiles/.

    python profiles/summarize.py r13            # reads gpurun_out/r13_tma.ncu-rep, r13_launches.csv, r13_bench.json
    python profiles/summarize.py r2z box_13_256_2 r2     # any other capture gpurun_out/<tag>_<which>.ncu-rep, round prefix r2

Writes profiles/<round>_<tag>_kernel.json (ncu --set full, one launch of the dominant kernel),
profiles/<round>_<tag>_launches.csv (every launch of the bench command with its device time),
profiles/<round>_<tag>_stalls.txt (warp-stall samples per reason and the top instructions) and
profiles/traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).
Needs `ncu` (present in the build container; the reports are read offline, no GPU).
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
ROUND = "r1"

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return v


def main():
    global ROUND
    tag = sys.argv[1] if len(sys.argv) > 1 else "r13"
    which = sys.argv[2] if len(sys.argv) > 2 else "tma"          # "tma" (bench.py capture) or any other capture name
    if len(sys.argv) > 3:
        ROUND = sys.argv[3]
    rep = os.path.join(OUT, f"{tag}_{which}.ncu-rep")
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {"source": f"gpurun_out/{tag}_{which}.ncu-rep (ncu --set full --clock-control none --import-source on, 1 launch)",
           "kernel": vals[hdr.index("Kernel Name")], "metrics": {}}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            rec["metrics"][k] = {"value": num(vals[i]), "unit": units[i]}
    for h, u, v in zip(hdr, units, vals):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            x = num(v)
            if isinstance(x, float) and x >= 0.01:
                rec["metrics"][h] = {"value": x, "unit": u}
    m = rec["metrics"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = m["dram__bytes_read.sum"]["value"] * scale[m["dram__bytes_read.sum"]["unit"]]
    wr = m["dram__bytes_write.sum"]["value"] * scale[m["dram__bytes_write.sum"]["unit"]]
    rec["dram_bytes_per_launch"] = rd + wr
    bench = os.path.join(OUT, f"{tag}_bench.json")
    sfx = "" if which == "tma" else f"_{which}"
    if which != "tma":
        plain = os.path.join(OUT, f"{tag}_plain3.log")
        for cand in (os.path.join(OUT, f"{tag}_plain_{which}.log"), os.path.join(OUT, f"{tag}_plain_{which.split('_', 1)[-1]}.log")):
            if os.path.isfile(cand):
                plain = cand
        if os.path.isfile(plain):
            rec["plain_run"] = open(plain).read().strip().splitlines()[-1]
    if which == "tma" and os.path.isfile(bench):
        line = [l for l in open(bench).read().splitlines() if l.startswith("{")][-1]
        b = json.loads(line)
        rec["bench_line"] = {k: b[k] for k in ("value", "unit", "ms_per_step", "roofline", "clocks", "gpu_launches")}
        algo_bytes = b["roofline"]["bytes_per_pair"] * b["config"]["patches_per_gpu"]
        rec["algorithmic_bytes_per_launch"] = algo_bytes
        rec["traffic_over_algorithmic"] = rec["dram_bytes_per_launch"] / algo_bytes
    json.dump(rec, open(os.path.join(ROOT, "profiles", f"{ROUND}_{tag}{sfx}_kernel.json"), "w"), indent=1)
    if which == "tma":
        json.dump({"dram_bytes_per_launch": rec["dram_bytes_per_launch"], "kernel": "degrade_tma_kernel<0, false>",
                   "source": f"profiles/{ROUND}_{tag}_kernel.json"}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)

    # launch list
    src = os.path.join(OUT, f"{tag}_launches.csv")
    if which == "tma" and os.path.isfile(src):
        lines = [l for l in open(src) if l.startswith('"')]
        open(os.path.join(ROOT, "profiles", f"{ROUND}_{tag}_launches.csv"), "w").writelines(lines)
        rd_ = list(csv.DictReader(io.StringIO("".join(lines))))
        tot = {}
        for r in rd_:
            name = r["Kernel Name"].split("(")[0]
            tot.setdefault(name, [0, 0.0])
            tot[name][0] += 1
            tot[name][1] += float(r["Metric Value"])
        allns = sum(v[1] for v in tot.values())
        rec["launch_shares"] = {k: {"launches": v[0], "ns": v[1], "share": v[1] / allns} for k, v in tot.items()}
        # the timed region of bench.py holds only libkmsr launches (torch kernels above are the untimed synthetic-data
        # setup; prepare_kernels runs once in the constructor): share of each kernel inside a step
        step = {k: v for k, v in tot.items() if "kmsr::" in k and "prepare_kernels" not in k and "ffma2_probe" not in k}
        sns = sum(v[1] for v in step.values())
        rec["step_shares"] = {k: {"launches": v[0], "ns_per_launch": v[1] / v[0], "share_of_step": v[1] / sns}
                              for k, v in step.items()}
        json.dump(rec, open(os.path.join(ROOT, "profiles", f"{ROUND}_{tag}_kernel.json"), "w"), indent=1)

    # stall samples
    srows = ncu_csv(rep, "source", ("--print-source", "sass"))
    h = srows[1]
    ix = {k: i for i, k in enumerate(h)}
    data = srows[2:]
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    tot_s = sum(int(r[ix["# Samples"]]) for r in data)
    agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
    base = int(data[0][0], 16)
    with open(os.path.join(ROOT, "profiles", f"{ROUND}_{tag}{sfx}_stalls.txt"), "w") as f:
        f.write(f"# warp-stall samples of {rec['kernel']} ({tag}); {tot_s} samples, {len(data)} SASS instructions\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            if v:
                f.write(f"{k:28s} {v:8d} {100.0 * v / tot_s:6.2f} %\n")
        f.write("\n# top 40 instructions by samples: offset, SASS, samples, %, dominant reasons\n")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:40]:
            st = sorted(((s, int(r[ix[s]])) for s in stalls if int(r[ix[s]]) > 0), key=lambda kv: -kv[1])[:3]
            f.write(f"{int(r[0], 16) - base:#07x} {r[1].strip()[:72]:72s} {int(r[ix['# Samples']]):6d} "
                    f"{100.0 * int(r[ix['# Samples']]) / tot_s:5.2f} {st}\n")
    print(json.dumps({k: rec[k] for k in rec if k != "metrics"}, indent=1)[:3000])


if __name__ == "__main__":
    main()
