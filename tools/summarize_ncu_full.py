"""Summarises `ncu -i REPORT.ncu-rep --page raw --csv` (a `--set full` capture) per launch: time, DRAM bytes, warp
instructions, occupancy, issue utilisation, pipe utilisation and the top stall reasons (warps per issue).

    ncu -i gpurun_out/X_full.ncu-rep --page raw --csv > /tmp/raw.csv && python tools/summarize_ncu_full.py /tmp/raw.csv
"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=None):
    i = col.get(name)
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def unit(name):
    i = col.get(name)
    return units[i] if i is not None else ""


def scaled(r, name):
    v = get(r, name)
    if v is None:
        return None
    u = unit(name)
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u, 1)


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
              or h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]
for r in data:
    name = re.sub(r"^void\s+", "", r[col["Kernel Name"]]).replace("<unnamed>::", "")
    t = scaled(r, "gpu__time_duration.sum")
    rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
    inst = get(r, "smsp__inst_executed.sum") or get(r, "sm__inst_executed.sum")
    occ = get(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    regs = get(r, "launch__registers_per_thread")
    issue = get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or get(r, "sm__inst_issued.avg.pct_of_peak_sustained_active")
    l2 = get(r, "lts__t_sector_hit_rate.pct")
    l1 = get(r, "l1tex__t_sector_hit_rate.pct")
    alu = get(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
    fma = get(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") or get(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")
    fp64 = get(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") or get(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
    lsu = get(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")
    bank = get(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
    dyn = get(r, "launch__shared_mem_per_block_dynamic")
    print(f"{name[:110]}  grid {r[col['Grid Size']]}")
    print(f"    time={t:.4g}ms  dram_rd={rd / 1e9:.4g}GB  dram_wr={wr / 1e9:.4g}GB  dram={(rd + wr) / t / 1e6:.0f}GB/s  warp_inst={inst:.4g}  "
          f"occupancy%={occ:.3g}  regs={regs:.0f}  smem_dyn={dyn}  issue%={issue:.4g}  L2hit%={l2:.4g}  L1hit%={l1:.4g}")
    print(f"    pipes%: alu={alu} fma={fma} fp64={fp64} lsu={lsu}  smem_bank_conflicts={bank}")
    st = []
    for h in stall_cols:
        if "issue_stalled" in h and h.startswith("smsp__average_warps_issue_stalled_") and "not_issued" not in h:
            v = get(r, h)
            if v:
                st.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    st.sort(reverse=True)
    print("    stalls: " + ", ".join(f"{n}={v:.2f}" for v, n in st[:6]))
    print()
