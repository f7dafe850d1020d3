"""Summarise an ncu report (raw page) / launch list into a few lines.  Usage:
   python tools/ncu_summary.py raw <file.ncu-rep>      python tools/ncu_summary.py launches <file.csv>
   python tools/ncu_summary.py src <file.ncu-rep>"""
import collections, csv, subprocess, sys

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[-1]
    d = {k: (x + " " + un if un and k != "Kernel Name" else x) for k, un, x in zip(h, u, v)}
    num = dict(zip(h, v))
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed.sum"]
    for k in keys:
        print("%-75s %s" % (k, d.get(k, "?")[:110]))
    st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(x) for k, x in num.items()
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and x not in ("", "n/a")}
    print("stalls/issue:", ", ".join("%s %.2f" % kv for kv in sorted(st.items(), key=lambda t: -t[1])[:7]))

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0].replace("void ", "").replace("hfq::dev::", ""), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print("%-40s n=%3d total %9.3f ms  share %5.1f%%" % (k, len(v), sum(v) / 1e6, 100 * sum(v) / tot))

def src(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]; ix = {n: i for i, n in enumerate(h)}; data = rows[2:]
    f = lambda r, n: float(r[ix[n]].replace(",", "") or 0) if r[ix[n]] not in ("", "n/a") else 0.0
    tot = sum(f(r, "# Samples") for r in data)
    byop = collections.Counter(); cnt = collections.Counter()
    for r in data:
        t = r[ix["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        byop[op] += f(r, "# Samples"); cnt[op] += f(r, "Instructions Executed")
    print("samples by opcode:", [(k, "%.1f%%" % (100 * v / tot)) for k, v in byop.most_common(8)])
    print("instructions by opcode:", [(k, "%.3g" % v) for k, v in cnt.most_common(10)])
    print("top lines:")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:8]:
        print("  %5.1f%%  %s" % (100 * f(r, "# Samples") / tot, r[ix["Source"]][:100]))

if __name__ == "__main__":
    {"raw": raw, "launches": launches, "src": src}[sys.argv[1]](sys.argv[2])
