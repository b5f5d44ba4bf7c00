"""ORACLE wrapper (test infrastructure): ctypes binding of oracle/build/libphysref.so, the float64
CPU restatement of the articulated control step.  See oracle/physics_ref.c for the scope note
("parity unpinned" versus Isaac Gym)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "build", "libphysref.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "physics_ref.c")):
        subprocess.check_call(["make", "-C", HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.phys_ref_control_step.restype = C.c_int
        _lib.phys_ref_control_step_ball.restype = C.c_int
        _lib.phys_ref_diagnostics.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def control_step(model, verts, cfg, root, dof_pos, dof_vel, pd_tar, ext_wrench=None, n_steps=1, ball=None, hits=None):
    """In-place on float64 arrays root[n,13], dof_pos[n,nd], dof_vel[n,nd] (and ball[n,13], hits[n] int32 when given);
    returns (rb[n,nb,13], contact[n,nb,3])."""
    n = root.shape[0]
    for a in (root, dof_pos, dof_vel, pd_tar):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    rb = np.zeros((n, model.nb, 13))
    cf = np.zeros((n, model.nb, 3))
    verts = np.ascontiguousarray(verts, np.float32)
    total = np.zeros(n, np.int32)
    step_hits = np.zeros(n, np.int32) if hits is not None else None
    for _ in range(n_steps):
        rc = lib().phys_ref_control_step_ball(C.byref(model), _p(verts), C.byref(cfg), C.c_int(n), _p(root), _p(dof_pos),
                                              _p(dof_vel), _p(pd_tar), _p(ext_wrench), _p(rb), _p(cf), _p(ball), _p(step_hits))
        if rc != 0:
            raise RuntimeError(f"phys_ref_control_step failed for env {-rc - 1}")
        if hits is not None:
            total += step_hits
    if hits is not None:
        hits[:] = total  # racket impacts accumulated over the n_steps control steps
    return rb, cf


_faces_keep = None


def set_hull_faces(planes, tris, ntris, tmax):
    """install the hull faces (abi.pack_faces) for the exact ball / body contact; process-wide, the arrays are kept alive here"""
    global _faces_keep
    planes, tris = np.ascontiguousarray(planes, np.float32), np.ascontiguousarray(tris, np.uint8)
    ntris = np.ascontiguousarray(ntris, np.int32)
    _faces_keep = (planes, tris, ntris)
    lib().phys_ref_set_hull_faces(_p(planes), _p(tris), _p(ntris), C.c_int(int(tmax)))


def clear_hull_faces():
    global _faces_keep
    _faces_keep = None
    lib().phys_ref_set_hull_faces(None, None, None, C.c_int(0))


def hull_sphere(model, verts, body, centre, radius):
    """the exact query alone: (hit, penetration, outward normal) for a sphere at `centre` (body frame)"""
    pen, nl = C.c_double(0.0), np.zeros(3)
    verts = np.ascontiguousarray(verts, np.float32)
    c = np.ascontiguousarray(centre, np.float64)
    hit = lib().phys_ref_hull_sphere(C.byref(model), _p(verts), C.c_int(int(body)), _p(c), C.c_double(float(radius)), C.byref(pen), _p(nl))
    return int(hit), pen.value, nl


def set_threads(n):
    """OpenMP threads of the env loop; returns the previous count"""
    return int(lib().phys_ref_set_threads(C.c_int(int(n))))


def diagnostics(model, cfg, root, dof_pos, dof_vel):
    out = np.zeros(12)
    lib().phys_ref_diagnostics(C.byref(model), C.byref(cfg), _p(np.ascontiguousarray(root)), _p(np.ascontiguousarray(dof_pos)),
                               _p(np.ascontiguousarray(dof_vel)), _p(out))
    return dict(mass=out[0], com=out[1:4], P=out[4:7], L=out[7:10], ke=out[10], pe=out[11])
