"""Summarise an .ncu-rep (run here, no GPU needed): key metrics per kernel + hottest source lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'launch__waves_per_multiprocessor', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'dram__sectors_read.sum']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:72s} {r[i]:>22s} {units[i]}")
    print('---')
cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                    capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(cs)))
cur, hdr, agg, nk = None, None, {}, 0
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name":
        nk += 1
        continue
    if r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].isdigit():
        try:
            inst = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
        except Exception:
            continue
        a = agg.setdefault((cur, int(r[0]), r[1][:88]), [0, 0]); a[0] += inst; a[1] += samp
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print("total warp-instructions", ti, "stall samples", ts)
print("--- top lines by stall samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1]/ts*100:5.1f}% samp {v[0]/ti*100:5.1f}% inst  {k[0]}:{k[1]}  {k[2]}")
print("--- top lines by instructions")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0]/ti*100:5.1f}% inst {v[1]/ts*100:5.1f}% samp  {k[0]}:{k[1]}  {k[2]}")
