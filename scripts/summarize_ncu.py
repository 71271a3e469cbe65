"""Summarise an ncu launch list (CSV) and a --set full report into a markdown file under profiles/."""
import collections
import csv
import subprocess
import sys


def launch_table(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
        k = row["Kernel Name"].split("(")[0][-60:]
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    out = ["| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
        out.append(f"| `{k}` | {cnt[k]} | {v / 1e6:.3f} | {100 * v / T:.2f} % |")
    out.append(f"| **total** | {sum(cnt.values())} | {T / 1e6:.3f} | 100 % |")
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]


def full_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    h, units, rows = r[0], r[1], r[2:]
    out = ["| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows))) + " |",
           "|---|---|" + "---:|" * len(rows)]
    for w in WANT:
        if w in h:
            i = h.index(w)
            out.append(f"| `{w}` | {units[i]} | " + " | ".join(row[i] for row in rows) + " |")
    stalls = []
    for i, c in enumerate(h):
        if "issue_stalled" in c and c.endswith("per_issue_active.ratio") or ("issue_stalled" in c and "warp_latency" in c and c.endswith(".ratio")):
            try:
                stalls.append((float(rows[-1][i]), c))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    out.append("")
    out.append("Top warp-stall reasons (last launch, cycles per issued instruction):")
    for v, c in stalls[:6]:
        out.append(f"* `{c.split('issue_stalled_')[-1].replace('_per_issue_active.ratio','')}`: {v:.2f}")
    return "\n".join(out)


if __name__ == "__main__":
    title, launches, rep, dst = sys.argv[1:5]
    extra = sys.argv[5] if len(sys.argv) > 5 else ""
    md = [f"# {title}", "", extra, "", "## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised: compare shares)", "",
          launch_table(launches), "", "## Dominant kernel (`ncu --set full --clock-control none --import-source on`)", "", full_metrics(rep), ""]
    open(dst, "w").write("\n".join(md))
    print(open(dst).read())
