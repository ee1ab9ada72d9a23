"""Summarise an ncu --set full report (.ncu-rep) as markdown: per-launch key metrics (raw page) and the
stall-reason mix + hottest SASS lines (source page).  Usage: python tools/ncu_summary.py rep.ncu-rep [title]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, title):
    print(f"## {title}\n\n`{rep.split('/')[-1]}` (ncu --set full --clock-control none --import-source on)\n")
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        print(f"### launch: `{name}`\n\n| metric | value |\n|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} {units[i]} |")
        print()
    src = page(rep, "source")
    # several kernels may follow each other; summarise the first
    h = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if not h:
        return
    hdr = src[h[0]]
    end = h[1] - 1 if len(h) > 1 else len(src)
    data = [r for r in src[h[0] + 1:end] if len(r) == len(hdr)]
    ix = {c: i for i, c in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
    stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    agg = sorted(((sum(int(r[ix[s]] or 0) for r in data), s) for s in stalls), reverse=True)
    print("### warp-state sampling (first launch)\n")
    print(", ".join(f"{s[6:]} {100 * n / tot:.1f}%" for n, s in agg[:8]) + "\n")
    print("| samples | share | SASS | top stall |\n|---|---|---|---|")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
        n = int(r[ix["# Samples"]])
        top = max(((int(r[ix[s]] or 0), s) for s in stalls))
        print(f"| {n} | {100 * n / tot:.1f}% | `{r[ix['Source']].strip()[:70]}` | {top[1][6:]} |")
    print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
