"""One cluster-fused aggregation launch over 14 c3 volumes (pass 1), for ncu."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laser_3d_reconstruction_b200 import _native as N
ctx = N.Context(0)
ms = C.c_float()
ctx.check(ctx.lib.l3d_sgbm_vgroup_time(ctx.h, 1152, 720, 128, 1944, 7776, 14, 1, 1, C.byref(ms)), "vgroup_time")
print(ms.value)
