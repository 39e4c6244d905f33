#!/usr/bin/env python
"""profiles/*.raw.csv (ncu --set full, --page raw --csv) -> profiles/ncu_summary.json, the per-kernel ceilings bench.py attaches to its roofline
object (`roofline.ncu`): L2 / L1 throughput, issue-slot utilisation, lanes per instruction, DRAM traffic, top stall reasons.
usage: python tools/make_ncu_summary.py C3 profiles/r2e_c3_k_*.raw.csv [C2 profiles/...]"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {  # kernel function -> bench.py's kernel key
    "k_trace_closest_persistent": "trace_closest(persistent)", "k_trace_closest_wide": "trace_closest", "k_trace_closest": "trace_closest(bvh2)", "k_light_bounce": "light_bounce", "k_camera_shade": "camera_shade",
    "k_camera_connect_deferred": "camera_connect", "k_camera_connect": "camera_connect", "k_shadow_resolve": "shadow_trace", "k_shadow_trace": "shadow_trace",
    "k_camera_merge_closure": "camera_merge_generic", "k_camera_merge_generic_batched": "camera_merge_generic", "k_camera_merge_coop": "camera_merge",
    "k_camera_continue": "camera_continue"}


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def to_bytes(v, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
    x = num(v)
    return None if (scale is None or x is None) else x * scale


def to_ms(v, unit):
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit)
    x = num(v)
    return None if (scale is None or x is None) else x * scale


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr, units, r = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}

    def get(k):
        if k in idx:
            return r[idx[k]], units[idx[k]]
        for h, i in idx.items():  # section-prefixed duplicates ("SM_B.TriageCompute.l1tex__t_sectors.sum")
            if h.endswith("." + k):
                return r[i], units[i]
        return None, None
    name = r[idx["Kernel Name"]]
    fn = re.sub(r"^void ", "", name).split("<")[0].split("(")[0]
    t = to_ms(*get("gpu__time_duration.sum"))
    out = {"kernel": fn, "file": os.path.basename(path), "time_ms": t, "registers": num(get("launch__registers_per_thread")[0]),
           "issue_active_pct": num(get("smsp__issue_active.avg.pct_of_peak_sustained_active")[0]),
           "lanes_per_inst": num(get("smsp__thread_inst_executed_per_inst_executed.ratio")[0]),
           "warps_active_pct": num(get("sm__warps_active.avg.pct_of_peak_sustained_active")[0]),
           "l1_hit_pct": num(get("l1tex__t_sector_hit_rate.pct")[0]), "l2_hit_pct": num(get("lts__t_sector_hit_rate.pct")[0])}
    rd, wr = to_bytes(*get("dram__bytes_read.sum")), to_bytes(*get("dram__bytes_write.sum"))
    l2, l1 = to_bytes(*get("lts__t_bytes.sum")), to_bytes(*get("l1tex__t_bytes.sum"))
    if l2 is None and num(get("lts__t_sectors.sum")[0]) is not None:
        l2 = num(get("lts__t_sectors.sum")[0]) * 32.0
    if l1 is None and num(get("l1tex__t_sectors.sum")[0]) is not None:
        l1 = num(get("l1tex__t_sectors.sum")[0]) * 32.0
    out["l2_throughput_pct_of_peak"] = num(get("lts__throughput.avg.pct_of_peak_sustained_elapsed")[0])
    out["l1_throughput_pct_of_peak"] = num(get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed")[0])
    if t:
        sec = t * 1e-3
        out.update({"dram_read_GB": None if rd is None else rd / 1e9, "dram_write_GB": None if wr is None else wr / 1e9,
                    "dram_GBps": None if (rd is None or wr is None) else (rd + wr) / sec / 1e9,
                    "l2_GBps": None if l2 is None else l2 / sec / 1e9, "l1_GBps": None if l1 is None else l1 / sec / 1e9})
    stalls = {}
    for k in idx:
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", k)
        if m and num(r[idx[k]]) is not None:
            stalls[m.group(1)] = num(r[idx[k]])
    out["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
    return fn, out


def main(argv):
    dst = os.path.join(ROOT, "profiles", "ncu_summary.json")
    data = json.load(open(dst)) if os.path.exists(dst) else {}
    workload = None
    for a in argv:
        if not a.endswith(".csv"):
            workload = a
            data.setdefault(workload, {})
            continue
        fn, s = summarise(a)
        data[workload][NAMES.get(fn, fn)] = s
    json.dump(data, open(dst, "w"), indent=1, sort_keys=True)
    print(dst, {w: sorted(v) for w, v in data.items()})


if __name__ == "__main__":
    main(sys.argv[1:])
