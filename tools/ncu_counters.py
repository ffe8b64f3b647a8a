#!/usr/bin/env python3
"""Turns `ncu --set full` captures (one frame per workload, tools/profile_frame.py) into profiles/ncu_counters.json, the
file bench.py reads its DRAM-true roofline from.  Per workload and launch group (k_closest = k_closest + its k_coop
launches, k_shadow = k_shadow + k_coop) it sums dram__bytes_read + dram__bytes_write over the launches of ONE frame and
takes the time-weighted mean of the hit rates / issue utilisation of the main kernel.  The file carries the hash of the
CUDA sources the captured binary was built from; bench.py ignores it when the sources have changed since.

  python tools/ncu_counters.py hw14_dragon_class=gpurun_out/x_prof_hw14.ncu-rep synthetic_10M=gpurun_out/x_prof_10m.ncu-rep
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

M = {
    "t": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "l2": "lts__t_sector_hit_rate.pct",
    "l1": "l1tex__t_sector_hit_rate.pct",
    "issue": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "ipc": "sm__inst_executed.avg.per_cycle_elapsed",
    "thr": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "occ": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lsb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1e-6)


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")
        rec = {"name": name, "ms": to_ms(d[M["t"]], u[M["t"]]),
               "dram": to_bytes(d[M["rd"]], u[M["rd"]]) + to_bytes(d[M["wr"]], u[M["wr"]])}
        for k in ("l2", "l1", "issue", "ipc", "thr", "occ", "dram_pct", "lsb"):
            try:
                rec[k] = float(d[M[k]].replace(",", ""))
            except Exception:
                rec[k] = None
        out.append(rec)
    return out


def group(launches, main, frames):
    """launch group of one frame: `main` kernel launches followed by their k_coop launches"""
    sel = []
    take_coop = False
    for l in launches:
        if (main + "<") in l["name"]:
            sel.append(l)
            take_coop = True
        elif "k_coop<" in l["name"] and take_coop:
            sel.append(l)  # (both passes of k_coop behind the main kernel)
        elif "k_coop<" not in l["name"]:
            take_coop = False
    if not sel:
        return None
    mains = [l for l in sel if (main + "<") in l["name"]]
    w = sum(l["ms"] for l in mains)

    def wmean(k):
        vals = [(l[k], l["ms"]) for l in mains if l[k] is not None]
        return sum(v * t for v, t in vals) / sum(t for _, t in vals) if vals else None

    return {"dram_bytes": sum(l["dram"] for l in sel) / frames, "kernel_ms": sum(l["ms"] for l in sel) / frames,
            "main_kernel_ms": w / frames, "launches_per_frame": len(sel) / frames,
            "l2_hit": wmean("l2"), "l1_hit": wmean("l1"), "issue_slot_util": wmean("issue"), "warp_exec_eff": wmean("thr"),
            "achieved_occupancy": wmean("occ"), "ipc_per_sm": wmean("ipc"), "dram_throughput_pct_of_peak": wmean("dram_pct"), "long_scoreboard": wmean("lsb")}


def main():
    out = {"_comment": "ncu --set full --clock-control none counters per frame and launch group; written by tools/ncu_counters.py. "
                       "dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum of the group's launches (main kernel + its k_coop).",
           "source_sha16": bench.source_sha16()}
    for arg in sys.argv[1:]:
        wl, rep = arg.split("=", 1)
        frames = 1
        if ":" in rep:
            rep, f = rep.rsplit(":", 1)
            frames = int(f)
        launches = load(rep)
        rec = {"file": os.path.basename(rep), "frames_captured": frames}
        for main_name in ("k_closest", "k_shadow"):
            g = group(launches, main_name, frames)
            if g:
                g["file"] = os.path.basename(rep)
                rec[main_name] = g
        out[wl] = rec
    p = os.path.join(ROOT, "profiles", "ncu_counters.json")
    json.dump(out, open(p, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
