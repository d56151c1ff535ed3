"""Attribute executed warp-instructions of one kernel in an .ncu-rep to source lines (inlined headers included).
usage: ncu_lines.py <rep> <kernel regex> <object .o> <mangled substring>"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, kre, obj, mangled = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = hdr.index("Instructions Executed"); ct = hdr.index("Thread Instructions Executed")
sass = [(r[1].strip(), int(r[ci]), int(r[ct])) for r in rows[hi + 1:] if len(r) > ct and r[ci].isdigit()]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l)
lines = []
cur = ("?", 0)
for l in dis[start + 1:]:
  if l.startswith("//---------------------") or (l.startswith(".text.") ):
    break
  m = re.search(r'//## File "([^"]+)", line (\d+)', l)
  if m:
    cur = (os.path.basename(m.group(1)), int(m.group(2)))
    continue
  m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
  if m:
    lines.append(cur)
print("sass rows", len(sass), "disasm instrs", len(lines))
agg = collections.Counter(); aggt = collections.Counter()
for (txt, n, t), loc in zip(sass, lines):
  agg[loc] += n; aggt[loc] += t
tot = sum(agg.values())
print("total warp instr", tot)
byfile = collections.Counter()
for (f, l), n in agg.items(): byfile[f] += n
print({k: "%.1f%%" % (100.0 * v / tot) for k, v in byfile.most_common()})
for (f, l), n in agg.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 45):
  print("%6.2f%%  %9d  lanes %4.1f  %s:%d" % (100.0 * n / tot, n, aggt[(f, l)] / max(n, 1), f, l))
