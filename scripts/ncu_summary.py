"""Turns ncu reports (read here, no GPU needed) into the markdown summaries under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/r02_launches_step.md "title"
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r02_ncu_xxx.md "title" [traffic_key]
"""
import collections, csv, io, json, os, re, subprocess, sys

METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__cluster_dim_x",
           "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    name = re.sub(r"ws::\(anonymous namespace\)::|ws::<unnamed>::|void ", "", name)
    return re.sub(r"\(.*", "", name)[:70]


def launches(path, out, title):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "second": 1e3}.get(unit, 1e-6)
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    lines = [f"# {title}", "", f"{sum(a[0] for a in agg.values())} launches, sum of kernel durations {tot:.2f} ms "
             "(per-launch times under ncu are cold-cache and serialised: compare SHARES).", "",
             "| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        lines.append(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.1f} % |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


def full(rep, out, title, traffic_key=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title}", ""]
    traffic = {}
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        lines += [f"## `{short(rec.get('Kernel Name', '?'))}`  grid {rec.get('Grid Size', '?')} block {rec.get('Block Size', '?')}", "",
                  "| metric | value |", "|---|---|"]
        dram = 0.0
        for h, u in zip(hdr, units):
            if h in METRICS or h.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"):
                lines.append(f"| {h} | {rec[h]} {u} |")
                if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    dram += float(rec[h].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        lines += ["", f"DRAM traffic of this launch: {dram / 1e6:.1f} MB", ""]
        traffic[short(rec.get("Kernel Name", "?"))] = dram
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)
    if traffic_key:
        tj = os.path.join(os.path.dirname(out), os.path.basename(out).replace(".md", "_traffic.json"))
        json.dump({traffic_key: max(traffic.values()) if traffic else None, "per_kernel": traffic}, open(tj, "w"), indent=1)
        print("wrote", tj)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None)
