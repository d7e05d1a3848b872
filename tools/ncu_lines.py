"""Per-source-line instruction / shared-wavefront / stall-sample attribution of one kernel from an .ncu-rep captured with
--import-source on (kernels built with -lineinfo):  python tools/ncu_lines.py report.ncu-rep [top] [inst|smp]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "inst"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, v = rows[0], rows[2]
print("kernel: %s" % v[h.index("Kernel Name")])
for w in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
          "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    if w in h:
        print("%-62s %s %s" % (w, v[h.index(w)], rows[1][h.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
lines = []
for n, hi in enumerate(heads):  # one section per source file
    fn = rows[hi - 2][1].split("/")[-1]
    h = rows[hi]
    ie, iw, ism = h.index("Instructions Executed"), h.index("L1 Wavefronts Shared"), h.index("# Samples")
    end = heads[n + 1] - 2 if n + 1 < len(heads) else len(rows)
    for r in rows[hi + 1:end]:
        if r and r[0] != "":
            try:
                lines.append((fn, int(r[0]), r[1], int(r[ie]), int(r[iw] or 0), int(r[ism] or 0)))
            except ValueError:
                pass
tot = sum(x[3] for x in lines)
smp = sum(x[5] for x in lines)
print("attributed instructions %d, stall samples %d; sorted by %s" % (tot, smp, key))
for fn, ln, s, n, w, sm in sorted(lines, key=lambda x: -(x[3] if key == "inst" else x[5]))[:top]:
    print("%-20s %4d %5.1f%% inst  wf %7.2fM  smp %4.1f%% | %s" % (fn[:20], ln, 100 * n / max(tot, 1), w / 1e6, 100 * sm / max(smp, 1), s.strip()[:90]))
