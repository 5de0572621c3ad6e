#!/usr/bin/env python
"""Digest an .ncu-rep (read here, no GPU needed) into the handful of numbers DESIGN.md and bench.py cite.

    python profiles/ncu_digest.py gpurun_out/<name>.ncu-rep [--json out.json] [--sass-mix]
"""
import argparse
import collections
import csv
import io
import json
import subprocess

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    # tensor pipe (tcgen05 MMAs run on the hmma sub-pipe) and the tensor-memory path
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
]
STALLS = "smsp__average_warps_issue_stalled_"
TAIL = "_per_issue_active.ratio"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--json")
    ap.add_argument("--sass-mix", action="store_true")
    a = ap.parse_args()
    hdr, units, launches = raw(a.rep)
    ix = {h: i for i, h in enumerate(hdr)}
    digest = []
    for row in launches:
        d = {"kernel": row[ix["Kernel Name"]]}
        for k in KEYS:
            if k in ix:
                d[k] = row[ix[k]] + (" " + units[ix[k]] if units[ix[k]] else "")
        stalls = {h[len(STALLS):-len(TAIL)]: float(row[i]) for h, i in ix.items()
                  if h.startswith(STALLS) and h.endswith(TAIL)}
        d["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        if "dram__bytes_read.sum" in ix:
            r, w = ix["dram__bytes_read.sum"], ix["dram__bytes_write.sum"]
            d["dram_bytes_per_launch"] = to_bytes(row[r], units[r]) + to_bytes(row[w], units[w])
        digest.append(d)
    for d in digest:
        for k, v in d.items():
            print(f"{k:80s} {v}")
        print()
    if a.sass_mix:
        out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "sass"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = rows[1]
        src, ex = h.index("Source"), h.index("Instructions Executed")
        mix, tot = collections.Counter(), 0
        for r in rows[2:]:
            try:
                n = int(r[ex])
            except Exception:
                continue
            t = r[src].split()
            op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "")).split(".")[0]
            mix[op] += n
            tot += n
        print("SASS mix (warp instructions executed):", tot)
        for op, n in mix.most_common(14):
            print(f"  {op:10s} {n:12d} {100 * n / tot:5.1f}%")
        digest[0]["sass_mix_pct"] = {op: round(100 * n / tot, 1) for op, n in mix.most_common(14)}
    if a.json:
        json.dump(digest[0] if len(digest) == 1 else digest, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
