"""Join an `ncu --page source --csv` dump (SASS rows) with `nvdisasm -g` line info of the same kernel and aggregate
stall samples / executed instructions per source line and per phase (line ranges of compose_coop.cu).

    python tools/ncu_source_by_line.py <source.csv> <kernel.sass> [top]
"""
import csv, re, sys, collections
src_csv, sass, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
lines = []
cur = ("?", 0)
for l in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
assert len(lines) == len(data), (len(lines), len(data))
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
keys = ["# Samples", "Instructions Executed", "stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_membar",
        "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_math", "stall_lg", "stall_mio", "stall_no_inst"]
agg = collections.defaultdict(lambda: collections.Counter())
phase = collections.defaultdict(lambda: collections.Counter())
PH = [(0, 292, "init"), (293, 310, "wave top / items wait"), (311, 466, "A1 match"), (467, 492, "A1 end / arcs wait"), (493, 573, "B emit"),
      (574, 604, "C rank"), (605, 660, "D resolve+setup"), (661, 10**9, "epilogue")]
last_main = 0
for (fn, ln), r in zip(lines, data):
    if fn == "compose_coop.cu": last_main = ln
    for k in keys: agg[(fn, ln)][k] += f(r, k)
    ph = next(n for a, b, n in PH if a <= last_main <= b)
    for k in keys: phase[ph][k] += f(r, k)
tot = sum(v["# Samples"] for v in agg.values()); toti = sum(v["Instructions Executed"] for v in agg.values())
print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
print("== by phase (attributed by the enclosing compose_coop.cu line)")
for a, b, n in PH:
    v = phase[n]
    if not v["# Samples"]: continue
    print(f"{n:24s} samples {v['# Samples']/tot*100:5.1f}%  inst {v['Instructions Executed']/toti*100:5.1f}%  long_sb {v['stall_long_sb']/tot*100:5.1f}%  "
          f"barrier {v['stall_barrier']/tot*100:4.1f}%  short_sb {v['stall_short_sb']/tot*100:4.1f}%  wait {v['stall_wait']/tot*100:4.1f}%  membar {v['stall_membar']/tot*100:4.1f}%  "
          f"branch {v['stall_branch_resolving']/tot*100:4.1f}%  sel+notsel {(v['stall_selected']+v['stall_not_selected'])/tot*100:4.1f}%")
print("== top lines")
for (fn, ln), v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    print(f"{fn}:{ln:<5d} samples {v['# Samples']/tot*100:5.1f}%  inst {v['Instructions Executed']/toti*100:5.1f}%  long_sb {v['stall_long_sb']/tot*100:5.1f}%  "
          f"barrier {v['stall_barrier']/tot*100:4.1f}%  short_sb {v['stall_short_sb']/tot*100:4.1f}%  membar {v['stall_membar']/tot*100:4.1f}%")
