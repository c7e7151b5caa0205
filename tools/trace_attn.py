"""Cycle-level event trace of one CTA of the tcgen05 attention kernel (debug build with -DLS_TC_TRACE).
   python tools/trace_attn.py build          (here: cross-compiles tools/ubench/liblongspec_trace.so)
   python tools/trace_attn.py [N=131072]     (on the GPU box)
Events per step t: 0 MMA warp saw P_FULL(t) | 1 MMA warp issued PV(t)+QK(t+2) | 2 softmax saw S_FULL(t) |
3 S loaded | 4 exp done | 5 P stored + arrived | 6 PV(t) issued | 7 K tile of QK(t+2) seen ready."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "tools", "ubench", "liblongspec_trace.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    from longspec_b200.build import NVCC_FLAGS, sources
    cmd = ["/usr/local/cuda/bin/nvcc"] + NVCC_FLAGS + ["-DLS_TC_TRACE", "-o", LIB] + sources() + ["-lcuda"]
    subprocess.run(cmd, check=True)
    print("built", LIB)
    sys.exit(0)

import ctypes as C  # noqa: E402

import torch  # noqa: E402

from longspec_b200 import _lib  # noqa: E402

_lib.LIB_PATH = LIB
lib = _lib.load()
lib.ls_debug_set_trace.restype = C.c_int
lib.ls_debug_set_trace.argtypes = [C.c_void_p]
lib.ls_debug_set_flags.restype = C.c_int
lib.ls_debug_set_flags.argtypes = [C.c_int]
assert lib.ls_debug_set_flags(int(os.environ.get("DBG", 0)) | (int(os.environ.get("BX", 0)) << 16)) == 0   # BX = traced blockIdx.x
buf = torch.zeros(8 * 512, dtype=torch.int64, device="cuda")
assert lib.ls_debug_set_trace(C.c_void_p(buf.data_ptr())) == 0
sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["131072"]) + ["2"]
exec(open(os.path.join(ROOT, "tools", "prof_attn.py")).read())
torch.cuda.synchronize()
ev = buf.cpu().view(8, 512)
t0 = int(ev[2, 0])
lo, hi = 60, 60 + int(os.environ.get('ROWS', 12))
print("step  Pfull_seen  issued | Sfull_seen  S_loaded  exp_done  P_arrived   (cycles since first S_FULL; steps", lo, "..", hi, ")")
for t in range(lo, hi):
    print(f"{t:4d}  " + "  ".join(f"{int(ev[e, t]) - t0:9d}" for e in range(6)))
import numpy as np  # noqa: E402
e = ev.numpy().astype(np.int64)
entry = int(e[7, 511])
nz = e[2][e[2] != 0]
if entry and nz.size:
    print(f"CTA bx={os.environ.get('BX', 0)}: kernel entry -> first S_FULL seen {int(nz[0]) - entry} cycles")
one_row_tile = os.environ.get("LONGSPEC_TC1", "1") != "0"        # attn_tc1_kernel: events 6 / 7 belong to the QK issuer warp
valid = int((e[0] != 0).sum())
sl = slice(20, max(valid - 10, 22))
print(f"MMA warp (per PAIR of steps, {valid} pairs traced):")
print("  mean pair period            :", float(np.diff(e[0, sl]).mean()))
print("  P_FULL seen -> pair issued  :", float((e[1, sl] - e[0, sl]).mean()))
print("  issued -> next P_FULL seen  :", float((e[0, sl.start + 1:sl.stop + 1] - e[1, sl]).mean()))
if int((e[6] != 0).sum()) > 30 and not one_row_tile:
    print("  P_FULL seen -> PV issued    :", float((e[6, sl] - e[0, sl]).mean()))
    print("  PV issued -> K tile ready   :", float((e[7, sl] - e[6, sl]).mean()))
    print("  K tile ready -> QK issued   :", float((e[1, sl] - e[7, sl]).mean()))
vs = int((e[2] != 0).sum())
ss = slice(40, max(vs - 20, 42))
if vs > 60:
    print(f"softmax warpgroup 0/1 (per step, {vs} steps traced):")
    print("  S_FULL seen -> S loaded     :", float((e[3, ss] - e[2, ss]).mean()))
    print("  S loaded -> exp done        :", float((e[4, ss] - e[3, ss]).mean()))
    print("  exp done -> P arrived       :", float((e[5, ss] - e[4, ss]).mean()))
    print("  P arrived -> next S_FULL seen (same WG):", float((e[2, 42:vs - 18] - e[5, 40:vs - 20]).mean()))

# one-row-tile kernel (LONGSPEC_TC1 != 0): events 7 / 6 = QK issuer saw its barriers / issued QK(q); 0 / 1 = PV issuer saw P_FULL(p) / issued PV(p)
if one_row_tile and int((e[6] != 0).sum()) > 60:
    n = int((e[6] != 0).sum())
    s2 = slice(20, n - 10)
    print(f"two-issuer view ({n} pairs): QK period {float(np.diff(e[7, s2]).mean()):.0f}, QK barriers seen -> issued {float((e[6, s2] - e[7, s2]).mean()):.0f}, "
          f"QK issued -> next barriers seen {float((e[7, 21:n - 9] - e[6, s2]).mean()):.0f} | PV period {float(np.diff(e[0, s2]).mean()):.0f}, "
          f"P_FULL seen -> PV issued {float((e[1, s2] - e[0, s2]).mean()):.0f}, PV issued -> next P_FULL seen {float((e[0, 21:n - 9] - e[1, s2]).mean()):.0f}")
    print(f"   QK(q) issued -> S_FULL(q) seen by softmax {float((e[2, s2] - e[6, s2]).mean()):.0f}; P(p) arrived -> PV(p) barriers seen {float((e[0, s2] - e[5, s2]).mean()):.0f}")
