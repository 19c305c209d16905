"""Inclusive per-source-line profile of one kernel: executed warp-instructions and stall samples of every SASS instruction
(ncu source page) attributed to the line of the OUTERMOST non-trivial frame of its inline stack (nvdisasm -gi), i.e. to
the statement of the kernel body (or of a chosen function) that the instruction was inlined into.

    python tools/sass_profile.py report.ncu-rep [kernel] [frame-function-substring] [cubin]

`cubin` (+ `cubin`.cu, the program text) is the file written by tools/prof_run.py / prof_n2.py under PBR_SAVE_CUBIN in the
profiled run; without it the cubin is rebuilt locally.  The join is by instruction index, verified opcode by opcode.
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
frame_fn = (sys.argv[3] or None) if len(sys.argv) > 3 else None
if len(sys.argv) > 4:
    cubin = sys.argv[4]
    src = open(cubin + ".cu").read().splitlines()
else:
    m = engine.Model(cases.readme_case("N2" if ("k4" in kernel or "k3" in kernel) else "N1"), device=None)
    d, s = C.c_void_p(), C.c_ulong()
    engine.lib().pbr_model_cubin(m._ctx, C.byref(d), C.byref(s))
    cubin = os.path.join(tempfile.mkdtemp(), "m.cubin")
    open(cubin, "wb").write(C.string_at(d.value, s.value))
    src = m.source.splitlines()
sass = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout


def encl(l):
    for k in range(l - 1, 0, -1):
        t = src[k - 1]
        if t.startswith(('PBR_DEV', 'PBR_NOINLINE', 'extern "C"', 'template', '__device__', 'struct')):
            t2 = src[k] if t.startswith('template') else t
            mm = re.search(r'(\w+)\s*\(', t2)
            return mm.group(1) if mm else t2[:40]
    return '?'


fn, stack, insts = None, [], []
pending = []
for line in sass.splitlines():
    mm = re.match(r'\s*\.text\.(\S+):', line)
    if mm:
        fn = mm.group(1)
        continue
    mm = re.findall(r'line (\d+)', line) if line.lstrip().startswith('//## File') else None
    if mm:
        pending.append([int(x) for x in mm])
        continue
    mm = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if mm and fn == kernel:
        if pending:
            # first comment line = innermost frame with its "inlined at" chain; flatten innermost -> outermost
            stack = []
            for p in pending:
                for x in p:
                    if not stack or stack[-1] != x:
                        stack.append(x)
            pending = []
        insts.append((mm.group(2), list(stack)))
    elif mm:
        pending = []

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if r and r[0] == 'Address':
        hdr, start = r, i + 1
        break
ii, ti, si, so = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
data = [r for r in rows[start:] if len(r) > ii]
if len(data) != len(insts):       # ptxas is not bit-reproducible across processes: near-identical schedules, join stays approximate
    print(f"  WARNING: {len(data)} profiled vs {len(insts)} disassembled instructions; joining the common prefix by index")
bad = sum(1 for r, (op, _) in zip(data, insts) if not re.sub(r'^@!?U?P\d+\s+', '', r[so].strip()).startswith(op))
print(f"{kernel}: {len(insts)} SASS instructions, opcode mismatches in the join: {bad}")
inner, outer = collections.defaultdict(lambda: [0, 0, 0]), collections.defaultdict(lambda: [0, 0, 0])
tot = tots = 0
for r, (op, st) in zip(data, insts):
    n, t, sm = int(r[ii]), int(r[ti]), int(r[si])
    tot += n
    tots += sm
    k_in = encl(st[0]) if st else '?'
    # outermost frame inside frame_fn (default: the kernel body = last frame before the kernel's own line)
    key = None
    for x in reversed(st):
        f = encl(x)
        if frame_fn is None:
            if not f.startswith(kernel):
                key = x
                break
        elif frame_fn in f:
            key = x
            break
    for dct, k in ((inner, k_in), (outer, key)):
        dct[k][0] += n
        dct[k][1] += t
        dct[k][2] += sm
print(f"  {tot} warp-instructions, {tots} samples")
print("  exclusive, by innermost function:   instr%  samples%  lanes")
for k, v in sorted(inner.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"  {100*v[0]/tot:6.2f}  {100*v[2]/max(tots,1):7.2f}  {v[1]/max(v[0],1):6.1f}   {k}")
print("  inclusive, by statement of the outer frame:   instr%  samples%  lanes")
for k, v in sorted(outer.items(), key=lambda kv: -kv[1][0])[:40]:
    txt = src[k - 1].strip()[:105] if k else '?'
    print(f"  {100*v[0]/tot:6.2f}  {100*v[2]/max(tots,1):7.2f}  {v[1]/max(v[0],1):6.1f}   L{k}: {txt}")
