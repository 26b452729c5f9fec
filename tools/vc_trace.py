"""Phase timeline of the persistent solver kernel (WL_VC_TRACE=1): python tools/vc_trace.py [N]"""
import ctypes, os, sys
import numpy as np
os.environ["WL_VC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_tgv
from waterlily_b200 import Simulation
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sim = make_tgv(Simulation, N)
eng = sim._eng
for _ in range(3):
    eng.mom_step((0.0, 0.0, 0.0))
buf = (ctypes.c_ulonglong * 8192)()
n = ctypes.c_int(0)
fn = eng.cabi.lib.wl_debug_vc_trace
fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
fn(eng.h, buf, 4096, ctypes.byref(n))
a = np.array(buf[:2 * n.value], dtype=np.uint64).reshape(-1, 2)
t0 = int(a[0, 1])
prev = t0
per_level = {}
for tag, t in a:
    tag, t = int(tag), int(t)
    lvl, ph = divmod(tag, 1000)
    dt = (t - prev) / 1e3
    per_level.setdefault(lvl, [0.0, 0])
    per_level[lvl][0] += dt; per_level[lvl][1] += 1
    print(f"{(t - t0) / 1e3:9.1f} us  +{dt:7.1f}  level {lvl} phase {ph}")
    prev = t
print("per level (us, phases):", {k: (round(v[0], 1), v[1]) for k, v in sorted(per_level.items())})
