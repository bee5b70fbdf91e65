"""Summarise an ncu report: headline metrics + hottest SASS lines (by stall samples) of the first kernel in it."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print("%-70s %-8s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(int(r[isamp] or 0), int(r[iex] or 0), i, r[isrc].strip()) for i, r in enumerate(rows[2:])]
tot = sum(d[0] for d in data)
print("total samples %d, warp instructions %d" % (tot, sum(d[1] for d in data)))
for s, e, i, t in sorted(data, reverse=True)[:top]:
    print("%5d %5.1f%% %10d  #%d %s" % (s, 100.0 * s / max(tot, 1), e, i, t[:110]))
