"""Summarises ncu artefacts brought back in gpurun_out/ into text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1b.csv > profiles/r1_launches_summary.txt
  python tools/ncu_summary.py full gpurun_out/prof_tc_r1b.ncu-rep > profiles/r1_tc_conv_ncu.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] == "us" else v)
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list (gpu__time_duration.sum, --clock-control none), one resident step, source: {path}")
    print(f"# total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':48s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:48]:48s} {v[0]:8d} {v[1]:10.3f} {v[1] / tot:7.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = rows[0]
    units = rows[1]
    print(f"# ncu --set full --clock-control none, source: {path}")
    for n, r in enumerate(rows[2:]):
        print(f"## launch {n}")
        d = dict(zip(H, r))
        u = dict(zip(H, units))
        for k in KEYS:
            if k in d:
                print(f"{k:72s} {d[k]} {u.get(k, '')}")
        stalls = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for k, v in d.items()
                  if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v}
        tot = sum(stalls.values()) or 1.0
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        print("warp stall samples (top): " + ", ".join(f"{k} {v / tot:.2f}" for k, v in top))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
