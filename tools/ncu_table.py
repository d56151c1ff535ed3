"""Per-kernel table (duration, DRAM bytes, achieved GB/s vs the measured peak, L2 / L1 hit rates, IPC, lanes) from an .ncu-rep.
usage: ncu_table.py <rep> [peak GB/s]"""
import csv, subprocess, sys
rep = sys.argv[1]; peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6550.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; units = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
def g(r, k):
  try: return float(r[idx[k]].replace(",", ""))
  except Exception: return float("nan")
def tob(r, k): return g(r, k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[k]], 1)
print("kernel | grid x block | duration us | DRAM read+write MB | achieved DRAM GB/s | %% of %.0f GB/s | L2 hit %% | L1 hit %% | IPC/SM | active lanes | warp-instr" % peak)
for r in rows[2:]:
  name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
  dur = g(r, "gpu__time_duration.sum"); dur_us = dur / 1000 if units[idx["gpu__time_duration.sum"]] == "ns" else dur
  b = tob(r, "dram__bytes_read.sum") + tob(r, "dram__bytes_write.sum")
  gbs = b / (dur_us * 1e-6) / 1e9
  print("%s | %s x %s | %.1f | %.2f | %.1f | %.2f | %.1f | %.1f | %.2f | %.1f | %.0f" % (
      name, r[idx["launch__grid_size"]], r[idx["launch__block_size"]], dur_us, b / 1e6, gbs, 100 * gbs / peak, g(r, "lts__t_sector_hit_rate.pct"),
      g(r, "l1tex__t_sector_hit_rate.pct"), g(r, "sm__inst_executed.avg.per_cycle_elapsed"), g(r, "smsp__thread_inst_executed_per_inst_executed.ratio"), g(r, "smsp__inst_executed.sum")))
