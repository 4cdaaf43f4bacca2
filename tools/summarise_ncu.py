"""Summaries of ncu CSV logs (read here, without a GPU):
    python tools/summarise_ncu.py launches gpurun_out/launches.csv  > profiles/rN_ncu_launch_summary.csv
    python tools/summarise_ncu.py traffic  gpurun_out/gemm_traffic.csv > profiles/rN_gemm_traffic.json
    python tools/summarise_ncu.py full     gpurun_out/prof.csv       > profiles/rN_ncu_<kernel>.csv   (from `ncu -i x.ncu-rep --page raw --csv`)
"""
import csv
import json
import re
import sys
from collections import defaultdict


def rows(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name.replace("void ", "").strip()


def launches(path):
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows(path):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        k = short(r["Kernel Name"])
        tot[k] += ms
        cnt[k] += 1
    total = sum(tot.values())
    print("kernel,launches,total_ms,share")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"{k},{cnt[k]},{tot[k]:.3f},{tot[k] / total:.4f}")
    print(f"TOTAL,{sum(cnt.values())},{total:.3f},1.0")


def traffic(path):
    per = defaultdict(dict)
    for r in rows(path):
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3,
                 "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(unit, 1)
        per[r["ID"]][r["Metric Name"]] = v * scale
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in per.values())
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in per.values())
    ms = sum(d.get("gpu__time_duration.sum", 0) for d in per.values())
    n = len(per)
    print(json.dumps({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                                "-k regex:gemm_tc over one denoise step (tools/profile_step.py --steps 1)",
                      "launches": n, "dram_bytes_read_total": rd, "dram_bytes_write_total": wr,
                      "dram_bytes_per_launch": (rd + wr) / max(n, 1), "gemm_ms_total_under_ncu": ms}, indent=1))


KEEP = re.compile(r"dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum|sm__pipe_tensor_cycles_active.*pct|"
                  r"sm__inst_executed_pipe_(xu|fma|alu).*pct|smsp__issue_active.*pct|sm__warps_active.*pct|"
                  r"launch__registers_per_thread|launch__shared_mem_per_block_dynamic|lts__t_sector_hit_rate.pct|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active|sm__throughput.*pct|smsp__inst_executed.sum$|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|gpu__dram_throughput.*pct|sm__cycles_active.avg$|"
                  r"l1tex__m_xbar2l1tex_read_bytes.sum$|l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed|"
                  r"lts__throughput.avg.pct_of_peak_sustained_elapsed|launch__cluster_size|sm__cycles_elapsed.max$")


def full(path):
    rs = rows(path)
    # raw page: one row per kernel, metrics as columns (first data row after the units row)
    hdr = rs[0].keys()
    data = [r for r in rs if r.get("ID", "").strip().isdigit()]
    units = rs[0] if not rs[0].get("ID", "").strip().isdigit() else {}
    print("metric,unit,value")
    r = data[0]
    for k in ("Kernel Name", "Block Size", "Grid Size"):
        if k in r:
            print(f'{k},,"{r[k]}"')
    for k in sorted(hdr):
        if KEEP.search(k):
            print(f"{k},{units.get(k, '')},{r[k]}")


if __name__ == "__main__":
    {"launches": launches, "traffic": traffic, "full": full}[sys.argv[1]](sys.argv[2])
