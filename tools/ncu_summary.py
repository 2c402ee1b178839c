"""Summarise ncu outputs into small text files for profiles/ (run in the build container, no GPU needed).
  python tools/ncu_summary.py launches gpurun_out/r01_launches.csv > profiles/r01_launches_summary.txt
  python tools/ncu_summary.py full gpurun_out/r01_gemm_full.ncu-rep > profiles/r01_gemm_full_summary.txt"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
    # where a GEMM is not tensor-bound: the SM's data paths (round 1, f16f8 arithmetic)
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_req_cycles_active.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = {}
    for row in csv.DictReader(lines):
        name = row["Kernel Name"][:100]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += ms
    s = sum(v[1] for v in tot.values())
    print(f"# {path}: per-kernel device time (ncu, serialised, cold cache) — compare SHARES\n# launches  total_ms  share  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[0]:4d} {v[1]:9.3f} {100 * v[1] / s:5.1f}%  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: ncu --set full, selected metrics per captured launch")
    for row in rows[2:]:
        print(f"\n{row[idx['Kernel Name']][:110]}")
        for m in METRICS:
            if m in idx:
                print(f"    {m:70s} {row[idx[m]]:>14s} {units[idx[m]]}")


def traffic(path):
    """DRAM bytes (read + write) of the LAST captured launch of every distinct kernel, and their sum = bytes per step
    when the capture covers one step: the numbers profiles/ncu_traffic.json carries for bench.py's roofline."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    per = {}
    for row in rows[2:]:
        name = row[idx["Kernel Name"]][:120]
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(row[idx[m]].replace(",", "")) * scale.get(units[idx[m]], 1.0)
        per[name] = b
    print(json.dumps({"dram_bytes_per_kernel": per, "dram_bytes_per_step": sum(per.values())}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
