"""Turn ncu CSV output into the small summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches_bench  "title" "command"
    python tools/summarize_ncu.py full gpurun_out/full_rec_fwd.csv [more.csv ...]     (prints markdown tables to stdout)
"""
import csv, io, sys, collections

def read_ncu_csv(path):
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))

def short(name):
    name = name.replace("b2t::", "").replace("void ", "")
    return name.split("(")[0]

if sys.argv[1] == "launches":
    rows = read_ncu_csv(sys.argv[2])
    out, title, cmd = sys.argv[3], sys.argv[4], sys.argv[5]
    per = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        per.append((short(r["Kernel Name"]), ns))
    with open(out + ".csv", "w") as f:
        f.write("kernel,duration_ns\n")
        for k, ns in per:
            f.write(f'"{k}",{ns:.0f}\n')
    agg = collections.OrderedDict()
    for k, ns in per:
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ns
    tot = sum(a[1] for a in agg.values())
    with open(out + ".md", "w") as f:
        f.write(f"# {title}\n\nCommand: `{cmd}` (per-launch times are serialised and cold-cache: compare shares, not absolutes).\n\n")
        f.write(f"kernel launches captured: {len(per)}; sum of durations: {tot / 1e6:.3f} ms\n\n| time (us) | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[0]} | `{k}` |\n")
    print(open(out + ".md").read())
else:
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__cycles_active.avg", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "launch__cluster_size"]
    for path in sys.argv[2:]:
        rows = list(csv.reader(open(path, errors="replace")))          # `--page raw --csv`: header row, unit row, one row per launch
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
            print(f"## {path.split('/')[-1]} - launch {m['ID'][0]}: `{short(m['Kernel Name'][0])}`\n\n| metric | value |\n|---|---|")
            for w in want:
                hit = [k for k in m if k == w or k.endswith("." + w)]
                for k in hit[:1]:
                    print(f"| {k} | {m[k][0]} {m[k][1]} |")
            print()
