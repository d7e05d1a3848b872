"""Per-source-line instruction / shared-wavefront / stall-sample attribution of one kernel from an .ncu-rep captured with
--import-source on (kernels built with -lineinfo):  python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, v = rows[0], rows[2]
for w in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    if w in h:
        print("%-62s %s" % (w, v[h.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
ie, iw, ism = h.index("Instructions Executed"), h.index("L1 Wavefronts Shared"), h.index("# Samples")
lines = []
for r in rows[hi + 1:]:
    if r and r[0] != "":
        try:
            lines.append((int(r[0]), r[1], int(r[ie]), int(r[iw] or 0), int(r[ism] or 0)))
        except ValueError:
            pass
tot = sum(x[2] for x in lines)
smp = sum(x[4] for x in lines)
print("attributed instructions %d, samples %d" % (tot, smp))
for ln, s, n, w, sm in sorted(lines, key=lambda x: -x[2])[:top]:
    print("%4d %5.1f%% inst  wf %6.2fM  smp %4.1f%% | %s" % (ln, 100 * n / tot, w / 1e6, 100 * sm / max(smp, 1), s.strip()[:100]))
