"""A/B of K1 on the same box: time 20 back-to-back K1 launches (10 M slots) with a given build of the library (raw ctypes)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcreg_b200.scenes import make_corridor
path = sys.argv[1]
lib = C.CDLL(path)
vp, dp = C.c_void_p, C.POINTER(C.c_double)
lib.dcreg_create.argtypes = [C.c_int, C.POINTER(vp)]
lib.dcreg_set_source.argtypes = [vp, C.POINTER(C.c_float), C.c_int64, C.c_int]
lib.dcreg_set_target.argtypes = [vp, C.POINTER(C.c_float), C.c_int64, C.c_int, C.c_double]
lib.dcreg_find_planes.argtypes = [vp, dp, C.c_double, dp, C.POINTER(C.c_int64)]
lib.dcreg_freeze_planes_f32.argtypes = [vp]
lib.dcreg_time_reduce.argtypes = [vp, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
h = vp()
assert lib.dcreg_create(0, C.byref(h)) == 0
scene = make_corridor(10_000_000, seed=44, noise=0.002)
T = np.eye(4); T[:3, 3] = [0.004, 0.003, -0.002]
fp = scene.ctypes.data_as(C.POINTER(C.c_float))
assert lib.dcreg_set_target(h, fp, len(scene), 3, 0.05) == 0
assert lib.dcreg_set_source(h, fp, len(scene), 3) == 0
npt = C.c_int64(0)
assert lib.dcreg_find_planes(h, T.ctypes.data_as(dp), 0.05, None, C.byref(npt)) == 0
assert lib.dcreg_freeze_planes_f32(h) == 0
prt = np.ascontiguousarray(np.concatenate([T[:3, :3].reshape(-1), T[:3, 3]]))
ms = C.c_float(0)
for wd in (0, 1):
    lib.dcreg_time_reduce(h, 0, prt.ctypes.data_as(dp), wd, 3, 0, C.byref(ms))
for rep in range(3):
    out = []
    for wd in (0, 1):
        lib.dcreg_time_reduce(h, 0, prt.ctypes.data_as(dp), wd, 20, 0, C.byref(ms))
        out.append(ms.value * 1e3)
    print(os.path.basename(path), "K1 f32 planes: wd off %.2f us, wd on %.2f us" % tuple(out))
