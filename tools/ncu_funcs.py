"""Aggregate the per-source-line output of tools/ncu_lines.py by enclosing function of xs_query.cuh.
usage: ncu_funcs.py <ncu_lines output with many lines>"""
import re, sys, os, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "cross-section_b200", "csrc", "xs_query.cuh")).read().splitlines()
funcs = []
for i, l in enumerate(src, 1):
  m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:XS_HD|XS_D|__device__ __forceinline__|__device__)\s+[\w:<> \*&]+?\s+(\w+)\(", l)
  if m: funcs.append((i, m.group(1)))
def fn(line):
  name = "?"
  for s, n in funcs:
    if s <= line: name = n
    else: break
  return name
agg = collections.Counter(); lanes = collections.Counter()
for l in open(sys.argv[1]):
  m = re.match(r"\s*([\d.]+)%\s+(\d+)\s+lanes\s+([\d.]+)\s+(\S+):(\d+)", l)
  if not m: continue
  n, la, f, line = int(m.group(2)), float(m.group(3)), m.group(4), int(m.group(5))
  key = fn(line) if f == "xs_query.cuh" else f
  agg[key] += n; lanes[key] += n * la
tot = sum(agg.values())
print("warp-instructions by function (share, average active lanes):")
for k, v in agg.most_common(30):
  print("%6.2f%%  lanes %4.1f  %s" % (100 * v / tot, lanes[k] / v, k))
