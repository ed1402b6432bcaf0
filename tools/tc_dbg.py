import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.tc_perf2 as t
from realcamnet_b200 import _C
lib = _C.lib()
lib.rcn_tc_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
def prof(*a, **k):
    lib.rcn_tc_prof(None, 1)
    t.run(*a, **k)
    out = (ctypes.c_ulonglong * 16)()
    lib.rcn_tc_prof(out, 0)
    n = max(out[2], 1)
    print(f"   prof: tiles {out[2]}  wait {out[0]/n:.0f} cyc/tile  work {out[1]/n:.0f} cyc/tile; q0: pre-waitld {out[4]/n:.0f} post-waitld {out[5]/n:.0f} post-transpose {out[6]/n:.0f}; q1 post-transpose {out[8]/n:.0f}; q3 post-transpose {out[9]/n:.0f}")
prof(2048, 128, 128, 1, dbgs=(128,))
prof(2048, 128, 128, 1, dbgs=(128 + 113,))
