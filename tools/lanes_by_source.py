"""Per source-function lane utilisation and stall samples of one kernel from an .ncu-rep (needs -lineinfo):
joins ncu's SASS page (by instruction order) with nvdisasm's line table of the cached cubin compiled from the same model.

    python tools/lanes_by_source.py report.ncu-rep [kernel]
"""
import collections
import csv
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyremot_b200 import cases, engine  # noqa: E402

rep = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 else "pbr_k2_solve_n1"
m = engine.Model(cases.readme_case("N1"), device=None)  # N1 and N2 share one program
d, s = C.c_void_p(), C.c_ulong()
engine.lib().pbr_model_cubin(m._ctx, C.byref(d), C.byref(s))
tmp = tempfile.mkdtemp()
open(os.path.join(tmp, "m.cubin"), "wb").write(C.string_at(d.value, s.value))
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, "m.cubin")], capture_output=True, text=True).stdout
src = m.source.splitlines()
fn, cur, lines = None, None, []
for line in sass.splitlines():
    mm = re.match(r'\s*\.text\.(\S+):', line)
    if mm:
        fn = mm.group(1)
    mm = re.search(r'//## File "[^"]*", line (\d+)', line)
    if mm:
        cur = int(mm.group(1))
        continue
    mm = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(\S+)', line)
    if fn == kernel and mm:
        lines.append((int(mm.group(1), 16), cur))
off2line = dict(lines)


def encl(l):
    for k in range(l - 1, 0, -1):
        t = src[k - 1]
        if t.startswith(('PBR_DEV', 'PBR_NOINLINE', 'extern "C"', 'template')):
            return (src[k] if t.startswith('template') else t)[:70]
    return '?'


out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if r and r[0] == 'Address':
        hdr, start = r, i + 1
        break
ii, ti, si = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
byline = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[start:]:
    try:
        addr, n, t, sm = int(r[0], 16), int(r[ii]), int(r[ti]), int(r[si])
    except Exception:
        continue
    if base is None:
        base = addr
    ln = off2line.get(addr - base)
    key = encl(ln) if ln else '?'
    for dct, k in ((agg, key), (byline, ln)):
        dct[k][0] += n
        dct[k][1] += t
        dct[k][2] += sm
tot = sum(v[0] for v in agg.values())
tots = sum(v[2] for v in agg.values())
print(f"{kernel}: {tot} warp-instructions, {tots} samples")
print("  instr%  samples%  avg lanes  function")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:14]:
    print(f"  {100*v[0]/tot:6.2f}  {100*v[2]/tots:7.2f}  {v[1]/max(v[0],1):8.1f}   {k}")
print("  lines with < 12 active lanes, by wasted lane-slots:")
low = [(v[0]*32 - v[1], ln, v) for ln, v in byline.items() if ln and v[0] and v[1]/v[0] < 12]
for w, ln, v in sorted(low, reverse=True)[:14]:
    print(f"  {100*v[0]/tot:6.2f}% instr  {v[1]/v[0]:5.1f} lanes   L{ln}: {src[ln-1].strip()[:100]}")
print("  top source lines by stall samples:")
for ln, v in sorted(byline.items(), key=lambda kv: -kv[1][2])[:28]:
    if ln:
        print(f"  {100*v[2]/tots:6.2f}% samples {100*v[0]/tot:6.2f}% instr {v[1]/max(v[0],1):5.1f} lanes  L{ln}: {src[ln-1].strip()[:110]}")
print("  top source lines by executed warp-instructions:")
for ln, v in sorted(byline.items(), key=lambda kv: -kv[1][0])[:45]:
    if ln:
        print(f"  {100*v[0]/tot:6.2f}% instr {100*v[2]/tots:6.2f}% samples {v[1]/max(v[0],1):5.1f} lanes  L{ln}: {src[ln-1].strip()[:110]}")
