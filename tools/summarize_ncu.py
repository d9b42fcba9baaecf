"""Turn ncu CSV exports into the short text summaries kept under profiles/.
  launches:  python tools/summarize_ncu.py launches <launches.csv>
  raw:       python tools/summarize_ncu.py raw <raw.csv> [kernel-substring]"""
import collections, csv, sys

def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        agg.setdefault(r[ki][:90], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("kernel | launches | mean us | share of GPU time")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%s | %d | %.2f | %.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg"]

def raw(path, sub=""):
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    stall = [i for i, c in enumerate(h) if "pcsamp_warps_issue_stalled" in c and "not_issued" not in c]
    for r in rows[2:]:
        if len(r) < len(h) or sub not in r[h.index("Kernel Name")]:
            continue
        print("==", r[h.index("Kernel Name")][:100])
        for k in KEYS:
            if k in h:
                print("   %-62s %s %s" % (k, r[h.index(k)], u[h.index(k)]))
        st = sorted(((float(r[i].replace(",", "") or 0), h[i].split("stalled_")[1]) for i in stall), reverse=True)[:5]
        print("   top stall samples:", ", ".join("%s=%d" % (n, v) for v, n in st))

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
