"""Opcode histogram per kernel of the built library (evidence that the hot kernels are tcgen05 / TMA / TMEM code):
   python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt        (runs here: cuobjdump needs no GPU)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "longspec_b200", "liblongspec_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, per = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and cur:
        per[cur][m.group(1)] += 1
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:110]
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "USETMAXREG", "HMMA", "LDGSTS", "LDSM",
        "MUFU", "FFMA2", "FADD2", "FMNMX3", "F2FP", "BAR", "STG", "LDG", "STS", "LDS", "ATOMG", "RED", "MEMBAR"]
print("# cuobjdump -sass longspec_b200/liblongspec_b200.so : opcode histogram per kernel (python -m longspec_b200.build:\n"
      "# nvcc 12.9 -gencode arch=compute_100a,code=sm_100a).  tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR, tcgen05.ld/st =\n"
      "# LDTM/STTM, TMA = UTMALDG, mbarrier = SYNCS, setmaxnreg = USETMAXREG, mma.sync = HMMA, cp.async = LDGSTS, ex2 = MUFU,\n"
      "# packed fp32x2 = FFMA2/FADD2, peer / system-scope stores and fences = STG + MEMBAR in the *_push / *_wait kernels.\n")
tot = collections.Counter()
for fn, c in per.items():
    tot.update(c)
    print(f"{demangle(fn)}\n    {sum(c.values())} instructions: " + ", ".join(f"{k} {c[k]}" for k in keys if c[k]))
print("\n# whole library\n    " + ", ".join(f"{k} {tot[k]}" for k in keys if tot[k]))
