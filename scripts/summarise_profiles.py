"""Turn the scratch ncu output of scripts/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/summarise_profiles.py r01
"""
import csv, io, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list -> per-kernel table + one step in order
rows = list(csv.reader(open(os.path.join(G, tag + "_launches_cfg3.csv"))))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
L = []
for r in rows[start:]:
    if len(r) > iv:
        v = float(r[iv].replace(",", ""))
        v = v / 1000 if r[iu] == "ns" else (v if r[iu] in ("us", "usecond") else v * 1000)
        L.append((r[ik], v))
agg = collections.OrderedDict()
for k, v in L:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in L)
with open(os.path.join(P, tag + "_launches_cfg3_nhwc_summary.txt"), "w") as f:
    f.write("# ncu launch list: python bench.py --steps 2 --warmup 3 --no-cpu --no-graph   (cfg3, batch 4096, tf32 path)\n")
    f.write("# gpu__time_duration.sum, --clock-control none; cold-cache / serialised: compare SHARES.  %d launches, %.1f us\n" % (len(L), tot))
    f.write("%-120s %5s %10s %6s\n" % ("kernel", "calls", "total_us", "share"))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-120s %5d %10.1f %5.1f%%\n" % (k[:120], n, v, 100 * v / tot))
    # the last full step, in launch order
    opt = [i for i, (k, _) in enumerate(L) if "opt_kernel" in k]
    if len(opt) >= 2:
        a, b = opt[-2] + 1, opt[-1] + 1
        f.write("\n# one step in launch order (%d launches, %.1f us)\n" % (b - a, sum(v for _, v in L[a:b])))
        for k, v in L[a:b]:
            f.write("%9.1f us  %s\n" % (v, k[:130]))
os.replace(os.path.join(G, tag + "_launches_cfg3.csv"), os.path.join(P, tag + "_launches_cfg3_nhwc.csv"))

# ---- --set full captures -> key metrics per kernel
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]
traffic = {}
with open(os.path.join(P, tag + "_ncu_full_nhwc.tsv"), "w") as f:
    f.write("kernel\tmetric\tunit\tvalue\n")
    for fn in sorted(os.listdir(G)):
        if not (fn.startswith(tag + "_full_") and fn.endswith(".ncu-rep")):
            continue
        raw = subprocess.run(["ncu", "-i", os.path.join(G, fn), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(io.StringIO(raw)))
        if len(r) < 3:
            continue
        h, u, v = r[0], r[1], r[2]
        name = v[h.index("Kernel Name")] if "Kernel Name" in h else fn
        for m in want:
            if m in h:
                j = h.index(m)
                f.write("%s\t%s\t%s\t%s\n" % (name[:90], m, u[j], v[j]))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tb = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            j = h.index(m)
            tb += float(v[j].replace(",", "")) * scale.get(u[j], 1)
        fam = fn[len(tag + "_full_"):-len(".ncu-rep")]
        traffic[fam] = {"bytes": tb, "instance": name[:120]}
json.dump(traffic, open(os.path.join(P, tag + "_traffic.json"), "w"), indent=1)
# ---- bench line
src = os.path.join(G, tag + "_bench_cfg3.json")
if os.path.exists(src):
    line = open(src).read().strip().splitlines()[-1]
    json.loads(line)
    open(os.path.join(P, tag + "_bench_cfg3_nhwc_tf32.json"), "w").write(line + "\n")
for c in ("cfg1", "cfg2", "cfg4", "cfg5", "ref"):
    src = os.path.join(G, "%s_bench_%s.json" % (tag, c))
    if os.path.exists(src) and os.path.getsize(src):
        line = open(src).read().strip().splitlines()[-1]
        json.loads(line)
        open(os.path.join(P, "%s_bench_%s.json" % (tag, c)), "w").write(line + "\n")
# ---- LSTM (cfg4) launch list -> per-kernel table
src = os.path.join(G, tag + "_launches_cfg4.csv")
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[start:]:
        if len(r) > iv:
            v = float(r[iv].replace(",", ""))
            v = v / 1000 if r[iu] == "ns" else (v if r[iu] in ("us", "usecond") else v * 1000)
            a = agg.setdefault(r[ik], [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    with open(os.path.join(P, tag + "_launches_cfg4_summary.txt"), "w") as f:
        f.write("# ncu launch list: python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --no-graph  (LSTM T=64, batch 1024, tf32 path)\n")
        f.write("# gpu__time_duration.sum, --clock-control none, first 120 launches; cold-cache / serialised: compare SHARES.  %.1f us\n" % tot)
        f.write("%-120s %5s %10s %6s\n" % ("kernel", "calls", "total_us", "share"))
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-120s %5d %10.1f %5.1f%%\n" % (k[:120], n, v, 100 * v / tot))
print("ok")
