"""Host-side mirror of the reference controller's per-cycle interface, on top of the C ABI.

    reference (dogbot_controller/src)                         here
    --------------------------------------------------------  ----------------------------------------
    DOGCTRL::update(...) + estimate() + QP + tau              WbcBatch.cycle(inputs) -> tau, w, x, ...
        client/main.cpp:572-660, 692-725, 984-1127, 1163-1397
    OPT(30, 86, 82); setQ; setc; setL_stance; opt_stance      OPT(...).setQ(...).setc(...).setL_stance(...).opt_stance()
        lopt.h:5-36, lopt.cpp:4-154

Everything numeric happens inside libwbc_b200.so (hand-written sm_100a CUDA behind include/wbc_b200.h).
This module only marshals pointers: numpy arrays for host buffers, torch CUDA tensors (or raw device
pointers) for device-resident buffers.  There is no CPU fallback: loading fails loudly when the
library is missing, and every call fails when no B200-class GPU is present.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

MODE_STANCE, MODE_SWING_BR_FL, MODE_SWING_BL_FR = 0, 1, 2
HOST_PTRS, DEVICE_PTRS, NO_SYNC, FIFO_DISPATCH, HOST_SLAB, SAMPLED_TRAJ = 0, 1, 2, 4, 8, 16

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

IN_FIELDS = [("base_pos", 3), ("base_rot", 9), ("base_rpy", 3), ("base_vel", 6), ("q", 12), ("dq", 12),
             ("com_des_pos", 6), ("com_des_vel", 6), ("com_des_acc", 6),
             ("sw_des_pos", 6), ("sw_des_vel", 6), ("sw_des_acc", 6), ("foot_force", 12), ("terrain", 40)]
OUT_FIELDS = [("tau", 12), ("w", 6), ("x", 30), ("qp_obj", 1)]
DEBUG_FIELDS = [("M", 324), ("h", 18), ("g", 18), ("Jac_lin", 216), ("Jdqd_lin", 12), ("com", 3), ("com_vel", 3),
                ("Mcom_b", 36), ("Mcom_j", 144), ("hcom", 18), ("gcom", 18), ("Jcom_lin", 216), ("Jdqdcom_lin", 12),
                ("foot_pos", 12), ("foot_vel", 12), ("Fgrf", 12), ("Wcom_des", 6)]


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("kcom", "dcom", "q1_weight", "slack_weight", "mu", "tau_max", "joint_dt", "kp_sw", "kd_sw", "g_acc",
                 "obs_gain", "obs_dt")] + [("gravity", C.c_double * 3), ("qp_epsx", C.c_double), ("qp_rho", C.c_double),
                                           ("obs_gain2", C.c_double),
                                           ("qp_outerits", C.c_int), ("observer_enabled", C.c_int),
                                           ("fix_swing_rhs", C.c_int), ("qp_literal_kkt", C.c_int),
                                           ("hold_tau_on_failure", C.c_int), ("obs_order", C.c_int), ("obs_form", C.c_int)]


class _Inputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _ in IN_FIELDS] + [("mode", C.c_void_p), ("ld", C.c_long), ("obs_gain", C.c_void_p)]


class _Outputs(C.Structure):
    _fields_ = [("tau", C.c_void_p), ("w", C.c_void_p), ("x", C.c_void_p), ("qp_obj", C.c_void_p), ("status", C.c_void_p),
                ("qp_info", C.c_void_p), ("qp_flops", C.c_void_p), ("ld", C.c_long), ("w3", C.c_void_p)]


class _Debug(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _ in DEBUG_FIELDS] + [("ld", C.c_long)]


class _PlantState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("base_pos", "base_rot", "base_rpy", "base_vel", "q", "dq", "foot_force", "mode")]


class _Trajectory(C.Structure):
    _fields_ = [("nseg", C.c_int), ("durations", C.c_void_p), ("nodes", C.c_void_p), ("ld", C.c_long)]


TRAJ_FIELDS = ["com_des_pos", "com_des_vel", "com_des_acc", "sw_des_pos", "sw_des_vel", "sw_des_acc"]


class _TrajSamples(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in TRAJ_FIELDS] + [("ld", C.c_long)]


class WbcError(RuntimeError):
    pass


_lib = None

EXPORTS = ["wbc_default_params", "wbc_last_error", "wbc_version", "wbc_create", "wbc_destroy", "wbc_set_params",
           "wbc_set_observer_state", "wbc_get_observer_state", "wbc_set_observer_state2", "wbc_get_observer_state2", "wbc_cycle", "wbc_debug_update", "wbc_debug_qp_records", "wbc_qp_solve",
           "wbc_plant_step", "wbc_plant_dynamics_step", "wbc_last_timing", "wbc_last_solve_cycles", "wbc_last_launches", "wbc_solver_shape", "wbc_stage_profile", "wbc_host_alloc",
           "wbc_host_free",
           "wbc_set_trajectory", "wbc_sample_trajectory",
           "wbc_measure_dfma_peak"]


def lib_path():
    return _build.LIB


def load():
    """dlopen libwbc_b200.so (building it in-tree first if a compiler is at hand).  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("WBC_B200_LIB") or _build.LIB          # A/B experiments: a variant build of the same library
    if path == _build.LIB:
        # rebuild when the library is missing or older than its sources (a no-op otherwise); on a box without nvcc the
        # shipped library is used as it is, and a missing one is an error: there is no CPU fallback
        if _build.have_nvcc() or not os.path.exists(_build.LIB):
            _build.build()
    lib = C.CDLL(path)
    lib.wbc_default_params.argtypes = [C.POINTER(Params)]
    lib.wbc_last_error.restype = C.c_char_p
    lib.wbc_version.restype = C.c_char_p
    lib.wbc_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(Params)]
    lib.wbc_destroy.argtypes = [C.c_void_p]
    lib.wbc_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
    lib.wbc_set_observer_state.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_long]
    lib.wbc_get_observer_state.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_long]
    lib.wbc_debug_qp_records.argtypes = [C.c_void_p, C.c_int, _dp, C.POINTER(C.c_int)]
    lib.wbc_set_observer_state2.argtypes = [C.c_void_p, C.c_int, _dp, C.c_long]
    lib.wbc_get_observer_state2.argtypes = [C.c_void_p, C.c_int, _dp, C.c_long]
    lib.wbc_cycle.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Inputs), C.POINTER(_Outputs), C.c_void_p, C.c_uint]
    lib.wbc_debug_update.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Inputs), C.POINTER(_Debug), C.c_uint]
    lib.wbc_qp_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
    lib.wbc_plant_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p,
                                   C.c_uint]
    lib.wbc_plant_dynamics_step.argtypes = [C.c_void_p, C.c_int, C.POINTER(_PlantState), C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_uint]
    lib.wbc_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.wbc_last_solve_cycles.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong)]
    lib.wbc_last_launches.argtypes = [C.c_void_p]
    lib.wbc_solver_shape.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.wbc_stage_profile.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
    lib.wbc_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    lib.wbc_host_free.argtypes = [C.c_void_p]
    lib.wbc_set_trajectory.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Trajectory), C.c_void_p, C.c_uint]
    lib.wbc_sample_trajectory.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.POINTER(_TrajSamples), C.c_void_p, C.c_uint]
    lib.wbc_measure_dfma_peak.argtypes = [C.c_void_p, _dp]
    _lib = lib
    return lib


def default_params():
    p = Params()
    load().wbc_default_params(C.byref(p))
    return p


def _check(rc, what):
    if rc != 0:
        raise WbcError("%s failed (%d): %s" % (what, rc, load().wbc_last_error().decode()))


def _ptr(a):
    """numpy array -> host pointer; torch tensor -> data_ptr(); int -> itself; None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return int(a)


class WbcBatch:
    """One ctx = one GPU.  Holds the per-instance observer state across cycles (main.cpp:243, 721-724)."""

    def __init__(self, max_batch, device=0, params=None):
        self.lib = load()
        self.params = params or default_params()
        self.max_batch = int(max_batch)
        self.device = int(device)
        h = C.c_void_p()
        _check(self.lib.wbc_create(C.byref(h), self.device, self.max_batch, C.byref(self.params)), "wbc_create")
        self.h = h
        self._pinned = []
        self.fifo_dispatch = False      # True: WBC_FIFO_DISPATCH (index-order work queue instead of longest-first)

    def close(self):
        if getattr(self, "h", None):
            self.lib.wbc_destroy(self.h)
            self.h = None
        for p in getattr(self, "_pinned", []):
            self.lib.wbc_host_free(p)
        self._pinned = []

    __del__ = close

    def set_params(self, params):
        self.params = params
        _check(self.lib.wbc_set_params(self.h, C.byref(params)), "wbc_set_params")

    def set_observer_state(self, yd, yw):
        yd = np.ascontiguousarray(yd, dtype=np.float64)
        yw = np.ascontiguousarray(yw, dtype=np.float64)
        n = yd.shape[1]
        _check(self.lib.wbc_set_observer_state(self.h, n, yd.ctypes.data_as(_dp), yw.ctypes.data_as(_dp), n),
               "wbc_set_observer_state")

    def get_observer_state(self, n):
        yd, yw = np.zeros((6, n)), np.zeros((6, n))
        _check(self.lib.wbc_get_observer_state(self.h, n, yd.ctypes.data_as(_dp), yw.ctypes.data_as(_dp), n),
               "wbc_get_observer_state")
        return yd, yw

    def set_observer_state2(self, yg):
        """ygamma, the extra integrator of the second-order observer (params.obs_order == 2)."""
        yg = np.ascontiguousarray(yg, dtype=np.float64)
        n = yg.shape[1]
        _check(self.lib.wbc_set_observer_state2(self.h, n, yg.ctypes.data_as(_dp), n), "wbc_set_observer_state2")

    def get_observer_state2(self, n):
        yg = np.zeros((6, n))
        _check(self.lib.wbc_get_observer_state2(self.h, n, yg.ctypes.data_as(_dp), n), "wbc_get_observer_state2")
        return yg

    @staticmethod
    def _inputs_struct(sc, n, keep, skip=()):
        ins = _Inputs()
        ld = None
        for name, k in IN_FIELDS:
            a = None if name in skip else sc.get(name)
            if a is None:
                if name != "terrain" and name not in skip:
                    raise WbcError("missing input array '%s'" % name)
                setattr(ins, name, None)
                continue
            if isinstance(a, np.ndarray):
                a = np.ascontiguousarray(a, dtype=np.float64)
                keep.append(a)
            if tuple(a.shape)[0] != k:
                raise WbcError("input '%s' must have shape [%d, ld]" % (name, k))
            ld = int(a.shape[1]) if ld is None else ld
            if int(a.shape[1]) != ld:
                raise WbcError("all inputs must share the same leading dimension")
            setattr(ins, name, _ptr(a))
        m = sc["mode"]
        if isinstance(m, np.ndarray):
            m = np.ascontiguousarray(m, dtype=np.int32)
            keep.append(m)
        ins.mode = _ptr(m)
        ins.ld = ld
        g = sc.get("obs_gain")                       # optional per-instance observer gain [n] (config 5's sweep)
        if g is not None:
            g = np.ascontiguousarray(g, dtype=np.float64).reshape(-1)
            if g.shape[0] != ld:
                raise WbcError("obs_gain must have one entry per instance")
            keep.append(g)
        ins.obs_gain = _ptr(g)
        if n > ld:
            raise WbcError("n exceeds the arrays' leading dimension")
        return ins

    def pinned(self, shape, dtype=np.float64):
        """A numpy array in page-locked host memory (wbc_host_alloc): wbc_cycle copies such arrays to / from the device
        directly instead of through its bounce buffer.  Freed with the batch."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _check(self.lib.wbc_host_alloc(C.byref(p), max(nbytes, 8)), "wbc_host_alloc")
        self._pinned.append(p)
        buf = (C.c_char * max(nbytes, 8)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def pinned_copy(self, a):
        out = self.pinned(a.shape, a.dtype)
        out[...] = a
        return out

    def set_trajectory(self, traj):
        """Upload a plan (scenarios.make_trajectory layout: nseg, durations [4*nseg, n], nodes [4*(nseg+1)*6, n])."""
        d = np.ascontiguousarray(traj["durations"], dtype=np.float64)
        nd = np.ascontiguousarray(traj["nodes"], dtype=np.float64)
        n = d.shape[1]
        tr = _Trajectory(int(traj["nseg"]), d.ctypes.data, nd.ctypes.data, n)
        _check(self.lib.wbc_set_trajectory(self.h, n, C.byref(tr), None, HOST_PTRS), "wbc_set_trajectory")

    def sample_trajectory(self, n, t=None, t_all=0.0, fetch=False):
        """Sample the uploaded plan on the GPU at per-instance times t [n] (or t_all for everyone).  The samples stay on
        the device for cycle(..., sampled_traj=True); fetch=True also returns them as a dict of [6, n] host arrays."""
        tp = None
        if t is not None:
            t = np.ascontiguousarray(t, dtype=np.float64)
            tp = t.ctypes.data
        _check(self.lib.wbc_sample_trajectory(self.h, n, tp, float(t_all), None, None, HOST_PTRS), "wbc_sample_trajectory")
        if not fetch:
            return None
        out = {k: np.zeros((6, n)) for k in TRAJ_FIELDS}
        o = _TrajSamples(*[out[k].ctypes.data for k in TRAJ_FIELDS], n)
        _check(self.lib.wbc_sample_trajectory(self.h, n, tp, float(t_all), C.byref(o), None, HOST_PTRS), "wbc_sample_trajectory")
        return out

    def pinned_inputs(self, sc):
        """Copy a scenario's input arrays into ONE page-locked slab laid out in wbc_inputs field order, so that
        wbc_cycle moves them with a single host-to-device copy.  Returns a dict of views (other keys passed through)."""
        n = int(sc["mode"].shape[0])
        fields = [(name, k) for name, k in IN_FIELDS if sc.get(name) is not None]
        rows = sum(k for _, k in fields) + (1 if sc.get("obs_gain") is not None else 0)
        slab = self.pinned((rows, n))
        out, r = dict(sc), 0
        for name, k in fields:
            out[name] = slab[r:r + k]
            out[name][...] = sc[name]
            r += k
        if sc.get("obs_gain") is not None:
            out["obs_gain"] = slab[r]
            out["obs_gain"][...] = sc["obs_gain"]
        out["mode"] = self.pinned_copy(np.ascontiguousarray(sc["mode"], dtype=np.int32))
        out["_slab"] = True          # cycle() passes WBC_HOST_SLAB
        return out

    def cycle(self, sc, n=None, want=("x", "qp_obj", "status", "qp_info", "qp_flops"), out=None, sampled_traj=False):
        """One control cycle on HOST (numpy) SoA inputs; returns a dict of numpy arrays [k, n].  `out`: preallocated
        result arrays to fill (e.g. page-locked ones from pinned()); its keys decide what is returned."""
        keep = []
        n = int(sc["mode"].shape[0]) if n is None else int(n)
        ins = self._inputs_struct(sc, n, keep, skip=TRAJ_FIELDS if sampled_traj else ())
        if out is None:
            out = {"tau": np.zeros((12, n)), "w": np.zeros((6, n))}
            if "x" in want: out["x"] = np.zeros((30, n))
            if "qp_obj" in want: out["qp_obj"] = np.zeros(n)
            if "status" in want: out["status"] = np.zeros(n, dtype=np.int32)
            if "qp_info" in want: out["qp_info"] = np.zeros((8, n), dtype=np.int32)
            if "qp_flops" in want: out["qp_flops"] = np.zeros(n)
            if "w3" in want: out["w3"] = np.zeros((12, n))
        o = _Outputs()
        for k in ("tau", "w", "x", "qp_obj", "status", "qp_info", "qp_flops", "w3"):
            setattr(o, k, _ptr(out.get(k)))
        o.ld = max(n, 1)
        _check(self.lib.wbc_cycle(self.h, n, C.byref(ins), C.byref(o), None, HOST_PTRS | (FIFO_DISPATCH if self.fifo_dispatch else 0) | (HOST_SLAB if sc.get("_slab") else 0) | (SAMPLED_TRAJ if sampled_traj else 0)), "wbc_cycle")
        return out

    def cycle_device(self, dev_in, dev_out, n, ld, stream=None, sync=True, out_ld=None):
        """One control cycle on DEVICE-resident SoA buffers (dicts of torch CUDA tensors or raw pointers).
        `ld` is the leading dimension of the input arrays, `out_ld` that of the output arrays (default: the outputs' own
        second dimension when they are tensors, else `ld`)."""
        if out_ld is None:
            t = dev_out.get("tau")
            out_ld = int(t.stride(0)) if hasattr(t, "stride") and t.dim() == 2 else ld
        # element (k, i) of an array is read or written at k * ld + i: refuse tensors whose row stride is not the leading
        # dimension they are passed with, or that are narrower than the batch (out-of-bounds accesses otherwise)
        for d, want_ld in ((dev_in, ld), (dev_out, out_ld)):
            for k, v in d.items():
                if hasattr(v, "stride") and hasattr(v, "dim"):
                    if v.dim() == 2 and (int(v.shape[1]) < n or (int(v.shape[0]) > 1 and int(v.stride(0)) != want_ld)):
                        raise WbcError("cycle_device: array %r has shape %s / row stride %d, expected >= %d columns and leading dimension %d"
                                       % (k, tuple(v.shape), int(v.stride(0)), n, want_ld))
                    if v.dim() == 1 and int(v.shape[0]) < n:
                        raise WbcError("cycle_device: array %r has %d entries, batch is %d" % (k, int(v.shape[0]), n))
        ins = _Inputs()
        for name, _ in IN_FIELDS:
            setattr(ins, name, _ptr(dev_in.get(name)))
        ins.mode = _ptr(dev_in["mode"])
        ins.obs_gain = _ptr(dev_in.get("obs_gain"))
        ins.ld = ld
        o = _Outputs()
        for k in ("tau", "w", "x", "qp_obj", "status", "qp_info", "qp_flops", "w3"):
            setattr(o, k, _ptr(dev_out.get(k)))
        o.ld = out_ld
        flags = DEVICE_PTRS | (0 if sync else NO_SYNC) | (FIFO_DISPATCH if self.fifo_dispatch else 0)
        _check(self.lib.wbc_cycle(self.h, n, C.byref(ins), C.byref(o), stream, flags), "wbc_cycle")

    def qp_records(self, n):
        """The QP records of the last cycle [n, doubles per record] (wbc_debug_qp_records; validation only)."""
        k = C.c_int(0)
        _check(self.lib.wbc_debug_qp_records(self.h, 0, None, C.byref(k)), "wbc_debug_qp_records")
        out = np.zeros((n, k.value))
        _check(self.lib.wbc_debug_qp_records(self.h, n, out.ctypes.data_as(_dp), None), "wbc_debug_qp_records")
        return out

    def debug_update(self, sc, n=None):
        """update() intermediates (wbc_debug_update) for stage-by-stage validation."""
        keep = []
        n = int(sc["mode"].shape[0]) if n is None else int(n)
        ins = self._inputs_struct(sc, n, keep)
        d = _Debug()
        out = {}
        for name, k in DEBUG_FIELDS:
            out[name] = np.zeros((k, n))
            setattr(d, name, out[name].ctypes.data)
        d.ld = n
        _check(self.lib.wbc_debug_update(self.h, n, C.byref(ins), C.byref(d), HOST_PTRS), "wbc_debug_update")
        return out

    def qp_solve(self, Q, c, L, neq):
        """n dense QPs, instance-major numpy arrays Q [n,30,30], c [n,30], L [n,nrows,31] (the OPT operator)."""
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        L = np.ascontiguousarray(L, dtype=np.float64)
        n, nrows = Q.shape[0], L.shape[1]
        x = np.zeros((n, 30))
        status = np.zeros(n, dtype=np.int32)
        info = np.zeros((n, 8), dtype=np.int32)
        flops = np.zeros(n)
        _check(self.lib.wbc_qp_solve(self.h, n, Q.ctypes.data, c.ctypes.data, L.ctypes.data, nrows, int(neq), x.ctypes.data,
                                     status.ctypes.data, info.ctypes.data, flops.ctypes.data, None, HOST_PTRS), "wbc_qp_solve")
        return x, status, info, flops

    def plant_step(self, base_pos, base_vel, push, foot_force=None, x=None, n=None, ld=None, stream=None, sync=True):
        """Synthetic plant of BASELINE config 5 (wbc_plant_step): advances base_pos [3,n] / base_vel [6,n] in place from the
        momentum balance of the last cycle and the world push wrench `push` [6,n]; with foot_force [12,n] and the last QP
        solution x [30,n] the loop is closed through the commanded ground-reaction forces.  numpy arrays (host), or torch
        CUDA tensors / raw device pointers (then pass n and ld)."""
        host = isinstance(base_vel, np.ndarray)
        if host:
            for a in (base_pos, base_vel, foot_force):
                assert a is None or (a.flags.c_contiguous and a.dtype == np.float64)
            n = int(base_vel.shape[1]); ld = n
            push = np.ascontiguousarray(push, dtype=np.float64)
            x = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
            flags = HOST_PTRS
        else:
            flags = DEVICE_PTRS | (0 if sync else NO_SYNC)
        _check(self.lib.wbc_plant_step(self.h, int(n), _ptr(base_pos), _ptr(base_vel), _ptr(foot_force), _ptr(x), _ptr(push), int(ld),
                                       stream, flags), "wbc_plant_step")

    def plant_dynamics_step(self, state, tau, push, n=None, ld=None, substeps=5, gamma=100.0, diag=None, stream=None, sync=True):
        """Forward-dynamics plant (wbc_plant_dynamics_step): `state` is a dict with base_pos, base_rot, base_rpy, base_vel, q, dq,
        foot_force, mode (numpy arrays: host path, advanced in place; torch CUDA tensors: device path)."""
        host = isinstance(state["q"], np.ndarray)
        n = int(state["mode"].shape[0]) if n is None else int(n)
        ld = n if ld is None else int(ld)
        ps = _PlantState()
        for k in ("base_pos", "base_rot", "base_rpy", "base_vel", "q", "dq", "foot_force", "mode"):
            setattr(ps, k, _ptr(state[k]))
        flags = (HOST_PTRS if host else DEVICE_PTRS) | (0 if (sync or host) else NO_SYNC)
        _check(self.lib.wbc_plant_dynamics_step(self.h, n, C.byref(ps), _ptr(tau), _ptr(push), ld, int(substeps), float(gamma), _ptr(diag), stream, flags),
               "wbc_plant_dynamics_step")

    def last_timing(self):
        a, b = C.c_float(0), C.c_float(0)
        _check(self.lib.wbc_last_timing(self.h, C.byref(a), C.byref(b)), "wbc_last_timing")
        return a.value, b.value

    def last_solve_cycles(self, n):
        """Per-instance solve duration of the last cycle in SM clock cycles (numpy uint64 [n])."""
        out = np.zeros(n, dtype=np.uint64)
        _check(self.lib.wbc_last_solve_cycles(self.h, n, out.ctypes.data_as(C.POINTER(C.c_ulonglong))), "wbc_last_solve_cycles")
        return out

    def last_launches(self):
        return int(self.lib.wbc_last_launches(self.h))

    def solver_shape(self):
        """(resident solver CTAs per SM, shared-memory bytes per CTA, grid of the last cycle) of the persistent solver kernel."""
        a, b, g, st = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        _check(self.lib.wbc_solver_shape(self.h, C.byref(a), C.byref(b), C.byref(g), C.byref(st)), "wbc_solver_shape")
        self.last_solver_kernel = "wbc_solve_staged_kernel" if st.value else "wbc_solve_kernel"
        return a.value, b.value, g.value

    def stage_profile(self):
        """Per-warp counters of the last staged solver launch (ctx created with WBC_STAGE_PROF=1): array [warps, 12]."""
        _, _, g = self.solver_shape()
        a = np.zeros((g, 12), dtype=np.uint64)
        rows = self.lib.wbc_stage_profile(self.h, a.ctypes.data_as(C.POINTER(C.c_ulonglong)), g)
        _check(min(rows, 0), "wbc_stage_profile")
        return a[:rows]

    def measure_dfma_peak(self):
        v = C.c_double(0)
        _check(self.lib.wbc_measure_dfma_peak(self.h, C.byref(v)), "wbc_measure_dfma_peak")
        return v.value


class OPT:
    """Mirror of the reference's OPT class (lopt.h:5-36): same method names and argument meaning, one QP per call.

    OPT(30, 86, 82) as constructed at main.cpp:266.  setQ / setc / setL_stance / setL_swing take dense arrays
    (the reference takes Eigen matrices by value); opt_stance / opt_swing return x (the reference fills a
    caller-allocated VectorXd and silently leaves it untouched on failure, lopt.cpp:114-116 -- here a failure
    raises WbcError instead).
    """

    def __init__(self, control_variables=30, stance_constraint=86, swing_constraint=82, batch=None):
        if control_variables != 30 or stance_constraint != 86 or swing_constraint != 82:
            raise WbcError("OPT is specialised to the controller's shapes OPT(30, 86, 82) (main.cpp:266)")
        self._own = batch is None
        self._b = batch or WbcBatch(1)
        self._Q = self._c = self._Ls = self._Lw = None

    def setQ(self, Q_):
        self._Q = np.array(Q_, dtype=np.float64).reshape(1, 30, 30)

    def setc(self, c_):
        self._c = np.array(c_, dtype=np.float64).reshape(1, 30)

    def setL_stance(self, L_stance):
        self._Ls = np.array(L_stance, dtype=np.float64).reshape(1, 86, 31)

    def setL_swing(self, L_swing):
        self._Lw = np.array(L_swing, dtype=np.float64).reshape(1, 82, 31)

    def _solve(self, L, neq):
        if self._Q is None or self._c is None or L is None:
            raise WbcError("OPT: setQ, setc and setL_* must be called before opt_*")
        x, status, _, _ = self._b.qp_solve(self._Q, self._c, L, neq)
        if status[0] != 0:
            raise WbcError("OPT: QP solver failed with status %d" % status[0])
        return x[0]

    def opt_stance(self):
        return self._solve(self._Ls, 18)     # first 18 rows are equalities (lopt.cpp:40-47)

    def opt_swing(self):
        return self._solve(self._Lw, 12)     # first 12 rows are equalities (lopt.cpp:57-64)
