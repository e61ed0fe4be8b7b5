import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laser_3d_reconstruction_b200 import _native as N
ctx = N.Context(0)
N_el = 1152 * 720 * 128
for njobs in (1, 7, 14, 15, 16):
    ms = C.c_float()
    ctx.check(ctx.lib.l3d_sgbm_vgroup_time(ctx.h, 1152, 720, 128, 1944, 7776, njobs, 1, 3, C.byref(ms)), "vgroup_time")
    per = ms.value / njobs
    print("njobs %2d: %.3f ms per launch, %.3f ms per job, DRAM-equivalent 6N -> %.2f TB/s" % (njobs, ms.value, per, 6 * N_el * 2 / 2 / (per * 1e-3) / 1e12 if False else 3 * N_el * 2 / (per * 1e-3) / 1e12))
