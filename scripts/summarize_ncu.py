"""Turns the ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
  python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep      > profiles/rNN_<kernel>_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for r in rd:
        ns = float(r[iv].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1)
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("bsg::<unnamed>::", "")
        d[name][0] += 1
        d[name][1] += ns
        n += 1
    tot = sum(v[1] for v in d.values())
    print(f"# ncu launch list: {n} launches, {tot / 1e6:.1f} ms of kernel time (gpu__time_duration.sum, --clock-control none;")
    print("# cold-cache, serialised: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for name, (c, t) in sorted(d.items(), key=lambda kv: -kv[1][1]):
        if t / tot < 0.0005:
            continue
        print(f"| `{name[:80]}` | {c} | {t / 1e6:.2f} | {100 * t / tot:.1f}% | {t / c / 1e3:.1f} |")


FULL_KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none: {path}\n")
    for r in rows[2:]:
        print(f"## launch {r[hdr.index('ID')]}: `{r[hdr.index('Kernel Name')][:100]}`\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for k in FULL_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
