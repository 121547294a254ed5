"""Summarise ncu outputs into small tracked files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r01_prof.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    return name.split("(")[0][:90]


def launches(src, dst):
    text = open(src).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += ns; total += ns
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none`; times are cold-cache and serialised:"
                " compare SHARES, not absolutes.\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% |\n")
        f.write(f"| **total** | {sum(a[0] for a in agg.values())} | {total / 1e6:.3f} | 100% |\n")
    print("wrote", dst)


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write(f"## `{short(d.get('Kernel Name', '?'))}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            stalls = [(h, d[h]) for h in hdr if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")]
            stalls = sorted(stalls, key=lambda kv: -float(kv[1].replace(",", "") or 0))[:8]
            for h, v in stalls:
                f.write(f"| {h} | {v} | % |\n")
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
