"""Attribute the SASS instructions of one kernel to the source functions they were inlined from (needs -lineinfo).

    python tools/sass_by_function.py [kernel]      # README N1 model, compile-only (no GPU)
"""
import collections
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyremot_b200 import cases, engine  # noqa: E402

kernel = sys.argv[1] if len(sys.argv) > 1 else "pbr_k2_solve_n1"
mi = cases.readme_case("N2" if "n2" in kernel else "N1")
m = engine.Model(mi, device=None)
d, s = C.c_void_p(), C.c_ulong()
engine.lib().pbr_model_cubin(m._ctx, C.byref(d), C.byref(s))
tmp = tempfile.mkdtemp()
open(os.path.join(tmp, "m.cubin"), "wb").write(C.string_at(d.value, s.value))
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, "m.cubin")], capture_output=True, text=True).stdout
src = m.source.splitlines()
fn, cur, cnt, ops = None, None, collections.Counter(), collections.Counter()
for line in sass.splitlines():
    mm = re.match(r'\s*\.text\.(\S+):', line)
    if mm:
        fn = mm.group(1)
    mm = re.search(r'//## File "[^"]*", line (\d+)', line)
    if mm:
        cur = int(mm.group(1))
        continue
    mm = re.match(r'\s+/\*[0-9a-f]+\*/\s+(\S+)', line)
    if fn == kernel and mm:
        cnt[cur] += 1
        ops[mm.group(1).split('.')[0].lstrip('@!P0123456789 ')] += 1


def encl(l):
    for k in range(l - 1, 0, -1):
        t = src[k - 1]
        if t.startswith(('PBR_DEV', 'PBR_NOINLINE', 'extern "C"', 'template')):
            return (src[k] if t.startswith('template') else t)[:90]
    return '?'


reg = collections.Counter()
for l, c in cnt.items():
    reg[encl(l) if l else '?'] += c
print(kernel, "total SASS instructions", sum(cnt.values()))
for k, v in reg.most_common(16):
    print(f"{v:7d}  {k}")
print("top opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
