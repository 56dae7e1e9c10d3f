"""Turns gpurun_out/ ncu artefacts into small committed summaries under profiles/ (usage: summarize_profiles.py TAG)."""
import collections, csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def launches():
    f = os.path.join(G, "launches.csv")
    if not os.path.exists(f): return
    rows = list(csv.reader(open(f))); hdr = None; agg = collections.defaultdict(lambda: [0, 0.0]); n = 0
    for r in rows:
        if len(r) > 5 and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r)); name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
            v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
            ms = v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v
            agg[name][0] += 1; agg[name][1] += ms; n += 1
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, tag + "_launches_summary.txt"), "w") as o:
        o.write("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 1 --warmup 3 --no-cpu-baseline\n")
        o.write("# %d launches captured (cold-cache, serialised: compare SHARES), total %.3f ms\n" % (n, tot))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write("%-64s n=%5d %10.3f ms %5.1f%% avg %9.1f us\n" % (k[:64], v[0], v[1], 100 * v[1] / tot, 1000 * v[1] / v[0]))

def raw(rep, out, keys):
    f = os.path.join(G, rep)
    if not os.path.exists(f): return
    txt = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3: return
    hdr = rows[0]; idx = [i for i, h in enumerate(hdr) if any(re.search(k, h) for k in keys)]
    with open(os.path.join(P, out), "w") as o:
        o.write("# ncu --set full --clock-control none --import-source on (%s); units row: %s\n" % (rep, ""))
        for r in rows[2:]:
            o.write("\n".join("%-70s %s %s" % (hdr[i], r[i], rows[1][i]) for i in idx) + "\n----\n")

def source(rep, out, top=30):
    f = os.path.join(G, rep)
    if not os.path.exists(f): return
    txt = subprocess.run(["ncu", "-i", f, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines())); hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hi: return
    hdr = rows[hi[0]]; end = hi[1] - 1 if len(hi) > 1 else len(rows)
    body = [r for r in rows[hi[0] + 1:end] if len(r) == len(hdr)]; col = {h: i for i, h in enumerate(hdr)}
    I = lambda r, h: int(float(r[col[h]] or 0))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(I(r, "# Samples") for r in body)
    with open(os.path.join(P, out), "w") as o:
        o.write("# warp-state samples of the first captured launch (%s), %d samples\n" % (rep, tot))
        for k, v in sorted({h: sum(I(r, h) for r in body) for h in stalls}.items(), key=lambda kv: -kv[1])[:10]:
            o.write("%-28s %7d %5.1f%%\n" % (k, v, 100.0 * v / max(1, tot)))
        o.write("# hottest SASS lines\n")
        for r in sorted(body, key=lambda r: -I(r, "# Samples"))[:top]:
            o.write("%7d samples exec %10d  %s\n" % (I(r, "# Samples"), I(r, "Instructions Executed"), r[col["Source"]][:90]))

KEYS = [r"Kernel Name", r"Grid Size", r"gpu__time_duration.sum", r"dram__bytes_(read|write).sum$", r"launch__registers_per_thread",
        r"sm__warps_active.avg.pct_of_peak", r"smsp__issue_active.avg.pct", r"sm__inst_executed.sum$", r"l1tex__data_bank_conflicts_pipe_lsu.sum$",
        r"sm__pipe_fma_cycles_active.avg.pct", r"sm__throughput.avg.pct", r"gpu__dram_throughput.avg.pct", r"lts__t_bytes.sum$", r"sm__pipe_tensor"]
launches()
import glob
for k in sorted(os.path.basename(f)[5:-8] for f in glob.glob(os.path.join(G, "prof_*.ncu-rep"))):
    raw("prof_%s.ncu-rep" % k, "%s_%s_raw.txt" % (tag, k), KEYS)
    source("prof_%s.ncu-rep" % k, "%s_%s_stalls.txt" % (tag, k))
for f in ("bench.json",):
    if os.path.exists(os.path.join(G, f)):
        open(os.path.join(P, tag + "_" + f), "w").write(open(os.path.join(G, f)).read())
print(sorted(os.listdir(P)))
