"""profiles/ncu_current.json from an `ncu --set full` capture (.ncu-rep) of the bench workload: per kernel (averaged over
its launches in the capture) duration, DRAM bytes read + written, L2 hit rate, IPC, active lanes, warp instructions.
bench.py reads the file for roofline.traffic / issue_rate, so those figures always name the capture they came from.
usage: ncu_current.py <rep> <capture name> [queries per launch] > profiles/ncu_current.json   (no GPU needed)"""
import csv, json, subprocess, sys
rep, name = sys.argv[1], sys.argv[2]
queries = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
def g(r, k):
  try: return float(r[idx[k]].replace(",", ""))
  except Exception: return None
def tob(r, k):
  v = g(r, k)
  return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[k]], 1)
acc = {}
for r in rows[2:]:
  k = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
  dur = g(r, "gpu__time_duration.sum")
  dur_us = dur / 1000 if units[idx["gpu__time_duration.sum"]] == "ns" else dur
  rd, wr = tob(r, "dram__bytes_read.sum"), tob(r, "dram__bytes_write.sum")
  e = acc.setdefault(k, {"n": 0, "duration_us": 0.0, "dram_bytes": 0.0, "l2_hit_pct": 0.0, "ipc": 0.0, "active_lanes": 0.0, "warp_inst": 0.0, "issue_active_pct": 0.0})
  e["n"] += 1
  for key, val in (("duration_us", dur_us), ("dram_bytes", (rd or 0) + (wr or 0)), ("l2_hit_pct", g(r, "lts__t_sector_hit_rate.pct")),
                   ("ipc", g(r, "sm__inst_executed.avg.per_cycle_elapsed")), ("active_lanes", g(r, "smsp__thread_inst_executed_per_inst_executed.ratio")),
                   ("warp_inst", g(r, "smsp__inst_executed.sum")), ("issue_active_pct", g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"))):
    e[key] += val or 0.0
out = {"capture": name, "queries": queries, "kernels": {}}
for k, e in acc.items():
  n = e.pop("n")
  out["kernels"][k] = {"launches_in_capture": n, **{key: round(v / n, 3) for key, v in e.items()}}
print(json.dumps(out, indent=1))
