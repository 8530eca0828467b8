#!/usr/bin/env python
"""Summarise ncu artifacts brought back from the GPU box into small, committed text files.

    python tools/ncu_summary.py --launches gpurun_out/X_launches.csv --rep gpurun_out/X_prof.ncu-rep --out profiles/rNN

writes <out>_launches.md (per-kernel share of the device time of one bench command, from the
`--metrics gpu__time_duration.sum` pass) and <out>_kernels.md (key counters of each kernel captured
with `--set full`: duration, registers, occupancy, issue/FP64-pipe utilisation, DRAM bytes, stall mix)."""
import argparse
import collections
import csv
import subprocess

STALLS = ["short_scoreboard", "long_scoreboard", "wait", "math_pipe_throttle", "barrier", "branch_resolving",
          "no_instruction", "mio_throttle", "lg_throttle", "dispatch_stall", "not_selected"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, mi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) * scale
    ours = {k: v for k, v in agg.items() if "tvf" in k or "kernel" in k and "at::" not in k}
    # set-up kernels outside bench.py's timed step: the FP64 peak probe and the input generator
    setup = ("fp64_peak", "sweep_seeds")
    tot = sum(v[1] for k, v in ours.items() if not any(s in k for s in setup))
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
        f.write("source: %s\n\n| kernel | launches | total us | avg us | share of the timed step's kernels |\n|---|---:|---:|---:|---:|\n" % path)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            share = "%.3f" % (v[1] / tot) if k in ours and not any(s in k for s in setup) else "(set-up, untimed)" if k in ours else "-"
            f.write("| %s | %d | %.1f | %.2f | %s |\n" % (k[:70], v[0], v[1], v[1] / v[0], share))


def kernels(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    ix = {h: i for i, h in enumerate(H)}

    def g(r, name, default=""):
        return r[ix[name]] if name in ix else default

    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none: key counters per captured kernel\n\nsource: %s\n\n" % rep)
        for r in rows[2:]:
            f.write("## %s\n\n" % g(r, "Kernel Name").split("(")[0])
            f.write("| metric | value |\n|---|---:|\n")
            for label, name in [
                ("duration (%s)" % (rows[1][ix["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in ix else ""), "gpu__time_duration.sum"),
                ("grid / block", None),
                ("registers per thread", "launch__registers_per_thread"),
                ("achieved warps active (% of peak)", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                ("warp instructions executed", "smsp__inst_executed.sum"),
                ("issue slots busy (%)", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                ("FP64 pipe active (% of peak)", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                ("avg active threads per instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
                ("dram bytes read (%s)" % (rows[1][ix["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in ix else ""), "dram__bytes_read.sum"),
                ("dram bytes written (%s)" % (rows[1][ix["dram__bytes_write.sum"]] if "dram__bytes_write.sum" in ix else ""), "dram__bytes_write.sum"),
                ("dram throughput (% of peak)", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                ("shared-memory bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                ("SM cycles elapsed (max)", "sm__cycles_elapsed.max"),
            ]:
                if name is None:
                    f.write("| %s | %s x %s |\n" % (label, g(r, "launch__grid_size"), g(r, "launch__block_size")))
                else:
                    f.write("| %s | %s |\n" % (label, g(r, name)))
            f.write("\nstall mix (warps stalled per issue-active cycle):\n\n| reason | ratio |\n|---|---:|\n")
            for s in STALLS:
                f.write("| %s | %s |\n" % (s, g(r, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s)))
            f.write("\n")


def counters_json(rep, out, problems_per_launch):
    """Machine-readable per-kernel counters (first captured launch of each kernel) for bench.py's `traffic`."""
    import json
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(H)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def num(r, name):
        return float(r[ix[name]].replace(",", "")) if name in ix and r[ix[name]] not in ("", "n/a") else None

    out_d = {"source": rep, "problems_per_launch": problems_per_launch, "kernels": {}}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
        launched = name
        name = {"tft_stage2_dual_kernel": "tft_stage2_kernel", "tft_moments_kernel": "tft_stage1_kernel", "tft_moments_tma_kernel": "tft_stage1_kernel",
                "tft_stage1_solve_dual_kernel": "tft_stage1_solve_kernel"}.get(name, name)   # the library's profile slot (tvf_kernel_name)
        if name in out_d["kernels"]:
            continue
        rd = num(r, "dram__bytes_read.sum"); wr = num(r, "dram__bytes_write.sum")
        out_d["kernels"][name] = {
            "dram_bytes_per_launch": (rd * scale.get(U[ix["dram__bytes_read.sum"]], 1.0) + wr * scale.get(U[ix["dram__bytes_write.sum"]], 1.0)) if rd is not None else None,
            "duration_us": num(r, "gpu__time_duration.sum"),
            "fp64_pipe_active_pct": num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warp_instructions": num(r, "smsp__inst_executed.sum"),
            "registers": num(r, "launch__registers_per_thread"),
            "launched_as": launched,
        }
        # executed FP64 thread-instructions of this launch: <op>.sum.per_cycle_elapsed x smsp__cycles_elapsed.avg
        cyc = num(r, "smsp__cycles_elapsed.avg")
        ops = {}
        for op in ("dfma", "dmul", "dadd"):
            v = num(r, "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op)
            ops[op] = v * cyc if (v is not None and cyc is not None) else None
        k = out_d["kernels"][name]
        k["fp64_thread_inst_per_launch"] = ops
        if all(v is not None for v in ops.values()):
            k["fp64_flop_per_launch"] = 2.0 * ops["dfma"] + ops["dmul"] + ops["dadd"]
            k["achieved_occupancy_pct"] = num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    json.dump(out_d, open(out, "w"), indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches"); ap.add_argument("--rep"); ap.add_argument("--out", required=True)
    ap.add_argument("--problems-per-launch", type=int, default=65536)
    a = ap.parse_args()
    if a.launches:
        launches(a.launches, a.out + "_launches.md")
    if a.rep:
        kernels(a.rep, a.out + "_kernels.md")
        counters_json(a.rep, a.out + "_counters.json", a.problems_per_launch)
