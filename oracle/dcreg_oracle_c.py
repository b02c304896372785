"""ctypes binding of oracle/dcreg_oracle.c (the C/OpenMP CPU oracle).  TEST INFRASTRUCTURE ONLY - see the header
of dcreg_oracle.c.  Used by tests/ and by bench.py's cpu_baseline / --impl reference leg."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libdcreg_oracle.so")
_lib = None


class Params(C.Structure):
    _fields_ = [("search_radius", C.c_double), ("max_iterations", C.c_int), ("detection", C.c_int),
                ("handling", C.c_int), ("use_weight_derivative", C.c_int), ("conv_rot", C.c_double),
                ("conv_trans", C.c_double), ("cond_thresh", C.c_double), ("eig_thresh", C.c_double),
                ("kappa_target", C.c_double), ("pcg_tol", C.c_double), ("pcg_max_iter", C.c_int),
                ("fixed_iterations", C.c_int), ("std_reg_gamma", C.c_double), ("thread_mode", C.c_int),
                ("n_threads", C.c_int)]


class Iter(C.Structure):
    _fields_ = [("n_eff", C.c_int), ("n_pt", C.c_int), ("is_degenerate", C.c_int), ("pcg_iterations", C.c_int),
                ("mask", C.c_int * 6), ("rmse", C.c_double), ("fitness", C.c_double), ("objective", C.c_double),
                ("H", C.c_double * 36), ("g", C.c_double * 6), ("dx", C.c_double * 6), ("T", C.c_double * 16),
                ("eig_full", C.c_double * 6), ("lam_schur_rot", C.c_double * 3), ("lam_schur_trans", C.c_double * 3),
                ("cond_schur_rot", C.c_double), ("cond_schur_trans", C.c_double), ("cond_diag_rot", C.c_double),
                ("cond_diag_trans", C.c_double), ("cond_full", C.c_double), ("P", C.c_double * 36)]


def build():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "dcreg_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


def available() -> bool:
    try:
        load()
        return True
    except Exception:
        return False


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        build()
    # one OpenMP thread per physical core unless the caller says otherwise: on the 64-core / 128-thread GPU-box host,
    # 128 libgomp threads ran the all-cores mode 20-40x SLOWER than 64 (measured, tools/cpu_probe.py)
    if "OMP_NUM_THREADS" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 2) // 2))
    lib = C.CDLL(_LIB)
    lib.orc_scene_create.restype = C.c_void_p
    lib.orc_scene_create.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int]
    lib.orc_scene_destroy.argtypes = [C.c_void_p]
    lib.orc_icp_run.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                C.POINTER(Iter), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.orc_analyze_and_solve.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(Params), C.POINTER(Iter)]
    assert lib.orc_sizeof_iter() == C.sizeof(Iter) and lib.orc_sizeof_params() == C.sizeof(Params)
    _lib = lib
    return lib


def make_params(search_radius=1.0, max_iterations=30, detection=1, handling=3, use_weight_derivative=False,
                conv_rot=1e-5, conv_trans=1e-3, cond_thresh=10.0, eig_thresh=120.0, kappa_target=1.0, pcg_tol=1e-6,
                pcg_max_iter=10, fixed_iterations=False, std_reg_gamma=0.01, thread_mode=1, n_threads=0) -> Params:
    return Params(search_radius, max_iterations, detection, handling, int(use_weight_derivative), conv_rot, conv_trans,
                  cond_thresh, eig_thresh, kappa_target, pcg_tol, pcg_max_iter, int(fixed_iterations), std_reg_gamma,
                  thread_mode, int(n_threads))


class Scene:
    def __init__(self, src, tgt):
        self.lib = load()
        self.src = np.ascontiguousarray(src[:, :3], dtype=np.float32)
        self.tgt = np.ascontiguousarray(tgt[:, :3], dtype=np.float32)
        fp = C.POINTER(C.c_float)
        self.h = self.lib.orc_scene_create(self.src.ctypes.data_as(fp), len(self.src), self.tgt.ctypes.data_as(fp), len(self.tgt))

    def close(self):
        if self.h:
            self.lib.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def icp_run(self, prm: Params, T_init, want_log=True):
        T_init = np.ascontiguousarray(T_init, dtype=np.float64)
        T_out = np.empty((4, 4))
        cap = prm.max_iterations if want_log else 0
        logs = (Iter * max(cap, 1))()
        n_it = C.c_int(0); conv = C.c_int(0)
        dp = C.POINTER(C.c_double)
        st = self.lib.orc_icp_run(self.h, C.byref(prm), T_init.ctypes.data_as(dp), T_out.ctypes.data_as(dp),
                                  logs if want_log else None, cap, C.byref(n_it), C.byref(conv))
        nrec = min(n_it.value, cap) if st != 1 else max(0, min(n_it.value - 1, cap))
        return st, bool(conv.value), n_it.value, T_out, [logs[i] for i in range(nrec)]


def max_threads() -> int:
    return int(load().orc_max_threads())


def time_icp(src, tgt, T0, iters, thread_mode=1):
    """Seconds for `iters` fixed ICP iterations ("Ours", released defaults), kd-tree build excluded.
    Returns (seconds, threads_used)."""
    sc = Scene(src, tgt)
    prm = make_params(max_iterations=iters, fixed_iterations=True, kappa_target=10.0, use_weight_derivative=False,
                      thread_mode=thread_mode)
    t0 = time.perf_counter()
    sc.icp_run(prm, T0, want_log=False)
    dt = time.perf_counter() - t0
    sc.close()
    return dt, (8 if thread_mode == 0 else max_threads())
