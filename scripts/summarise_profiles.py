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

# ---- --set full capture of one whole step -> key metrics per kernel (launch order) + DRAM traffic per call
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
# kernel instance (template arguments) -> the C-ABI call of the cfg3 step at batch 4096 it serves (bench.py's call labels)
CALL_OF = [
    ("conv_c3_fwd_kernel<32, 0, 1>", "bf_conv_c3_forward_pool(4096,3,30,30,32,3,3,0,1)"),
    ("conv_igemm_kernel<64, 1, 0, 0, 1>", "bf_conv_nhwc_forward_pool_packed(4096,32,14,14,64,3,3,1)"),
    ("conv_igemm_kernel<128, 2, 0, 0, 1>", "bf_conv_nhwc_forward_pool_packed(4096,64,6,6,128,3,3,1)"),
    ("dense_head_kernel<16>", "bf_dense_head(4096,128,4,10,4,1,1,1)"),
    ("conv_df_kernel<4>", "bf_conv_nhwc_backward_filter_dbp(1184,4096,64,6,6,128,3,3)"),
    ("conv_igemm_kernel<64, 2, 0, 0, 0>", "bf_conv_nhwc_backward_data_packed(4096,64,6,6,128,3,3)"),
    ("conv_df_kernel<2>", "bf_conv_nhwc_backward_filter_dbp(1184,4096,32,14,14,64,3,3)"),
    ("conv_igemm_kernel<32, 1, 0, 1, 0>", "bf_conv_nhwc_backward_data_packed(4096,32,14,14,64,3,3)"),
    ("conv_c3_df_kernel<1, 27, 1>", "bf_conv_c3_backward_filter_unpool(4096,3,30,30,32,3,3)"),
]
traffic = {}
rep = os.path.join(G, tag + "_full_step.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, u = r[0], r[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    with open(os.path.join(P, tag + "_ncu_full_step.tsv"), "w") as f:
        f.write("# ncu --set full --clock-control none of ONE cfg3 training step (batch 4096, tf32 path, eager launches), launch order\n")
        f.write("order\tkernel\t" + "\t".join(want) + "\n")
        f.write("\t\t" + "\t".join(u[h.index(m)] if m in h else "" for m in want) + "\n")
        pool_seen = 0
        for i, v in enumerate(r[2:]):
            name = v[h.index("Kernel Name")]
            f.write("%d\t%s\t%s\n" % (i, name[:70], "\t".join(v[h.index(m)] if m in h else "" for m in want)))
            tb = sum(float(v[h.index(m)].replace(",", "")) * scale.get(u[h.index(m)], 1) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            dur = float(v[h.index("gpu__time_duration.sum")].replace(",", ""))
            dur = dur / 1000 if u[h.index("gpu__time_duration.sum")] in ("ns", "nsecond") else dur
            entry = {"bytes": tb, "instance": name[:120], "ncu_us": dur,
                     "tensor_pipe_pct": float(v[h.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")] or 0)}
            for inst, call in CALL_OF:
                if inst in name and call not in traffic:
                    traffic[call] = entry
            if "pool_nhwc_bwd_kernel" in name:       # two launches per step: 128-channel layer first, then the 64-channel one
                call = ("bf_pool_nhwc_backward_db(4096,128,4,4,2)", "bf_pool_nhwc_backward_db(4096,64,12,12,2)")[min(pool_seen, 1)]
                traffic.setdefault(call, entry)
                pool_seen += 1
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
