#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small, committed text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.md
    python tools/ncu_summary.py full gpurun_out/pscv_r1.ncu-rep profiles/r1_pscv_full.md [--json profiles/pscv_l2_traffic.json]

`launches`: per-kernel totals of the `--metrics gpu__time_duration.sum` pass (cold-cache, serialised: shares only).
`full`: the raw-page metrics of an `ncu --set full` capture that the roofline discussion in DESIGN.md uses.
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def launches(src, dst):
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    tot = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        ns = float(r["Metric Value"].replace(",", ""))
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += ns
    total = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): gpu__time_duration.sum per kernel, {len(rows)} launches, {total/1e6:.3f} ms\n\n")
        f.write("Cold-cache, serialised replay: compare SHARES, not absolutes.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns/1e3:.1f} | {100*ns/total:.1f}% |\n")
    print(open(dst).read())


def full(src, dst, json_dst=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, zip(units, r)))
            f.write(f"\n## {d['Kernel Name'][1]}  grid {d.get('Grid Size', ('', ''))[1]} block {d.get('Block Size', ('', ''))[1]}\n\n| metric | unit | value |\n|---|---|---:|\n")
            for k in hdr:
                if k in KEYS or any(k.endswith(x) for x in ("xu_realtime.avg.pct_of_peak_sustained_elapsed", "alu_realtime.avg.pct_of_peak_sustained_elapsed")):
                    f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            if json_dst:
                def val(k):
                    u, v = d[k]
                    v = float(v.replace(",", ""))
                    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
                json.dump({"kernel": d["Kernel Name"][1], "source": src,
                           "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                           "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                           "gpu_time_us_under_ncu": float(d["gpu__time_duration.sum"][1].replace(",", ""))}, open(json_dst, "w"), indent=1)
                json_dst = None
    print(open(dst).read())


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "launches":
        launches(src, dst)
    else:
        full(src, dst, sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None)
