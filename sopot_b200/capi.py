"""ctypes binding of libsopot_b200.so (include/sopot_b200.h) used by tests/ and bench.py.

This is plumbing, not the product: every method is a 1:1 call into the C ABI.  Arrays are either
numpy float64 host arrays (MEM_HOST: the library stages them) or DeviceArray objects allocated by
the library itself (MEM_DEVICE).  If the shared library is missing or no sm_100 GPU is visible the
constructor raises -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_int32, c_size_t, c_uint32, \
    c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SOPOT_B200_LIB lets tools/ load an experimental build of the same library (never a different backend)
LIB_PATH = os.environ.get("SOPOT_B200_LIB") or os.path.join(_HERE, "lib", "libsopot_b200.so")

MEM_DEVICE, MEM_HOST = 0, 1
FP_CONTRACT, FP_STRICT = 0, 1
FLAG_GRID_PER_STAGE = 1
FLAG_GRID_MARCH, FLAG_GRID_TILE = 4, 8
FLAG_UGRID_MARCH, FLAG_UGRID_GATHER = 16, 32
FLAG_UGRID_STAGE_PAIRS = 2
ABI_VERSION = 5
INTEGRATOR_RK4, INTEGRATOR_SYMPLECTIC_EULER, INTEGRATOR_VELOCITY_VERLET = 0, 1, 2      # sopot_b200_ugrid_integrate
         # SOPOT_B200_ABI_VERSION of include/sopot_b200.h this table was written against
ST_OK, ST_NONFINITE, ST_SINGULAR = 0, 1, 2

_dp = POINTER(c_double)


class SopotB200Error(RuntimeError):
    pass


class RK4Opts(ctypes.Structure):
    _fields_ = [("mem", c_uint32), ("fp_mode", c_uint32), ("store_every", c_size_t),
                ("broadcast", c_uint32), ("flags", c_uint32)]


class HoParams(ctypes.Structure):
    _fields_ = [(k, c_void_p) for k in ("m", "k", "c", "x0", "v0")]


class DpendParams(ctypes.Structure):
    _fields_ = [(k, c_void_p) for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")]


class CartpoleParams(ctypes.Structure):
    _fields_ = [(k, c_void_p) for k in ("mc", "m", "L", "g", "x", "th", "xd", "w", "force")]


class CartpolePlant(ctypes.Structure):
    _fields_ = [(k, c_void_p) for k in ("mc", "m", "L", "g")]


class Feedback(ctypes.Structure):
    _fields_ = [("K", c_void_p), ("x_ref", c_void_p), ("u_min", c_double), ("u_max", c_double), ("saturate", c_int)]


class Table2D(ctypes.Structure):
    _fields_ = [("rows", c_size_t), ("cols", c_size_t), ("row_vals", c_void_p), ("col_vals", c_void_p), ("data", c_void_p)]


AERO_TABLE_NAMES = ("cd_power_on", "cd_power_off", "cl_power_on", "cl_power_off", "cs_power_on", "cs_power_off", "cmq_yz")


class AeroTables(ctypes.Structure):
    _fields_ = [(k, Table2D) for k in AERO_TABLE_NAMES] + [("engine_on", c_int)]


class RocketParams(ctypes.Structure):
    _fields_ = [("elevation_deg", c_void_p), ("azimuth_deg", c_void_p), ("diameter", c_void_p),
                ("gravity", c_void_p),
                ("engine_rows", c_size_t), ("engine", c_void_p), ("mass_rows", c_size_t), ("mass", c_void_p),
                ("ix_rows", c_size_t), ("ix", c_void_p), ("iyz_rows", c_size_t), ("iyz", c_void_p),
                ("aero", POINTER(AeroTables))]


class Dispersion(ctypes.Structure):
    _fields_ = [(k, c_double) for k in ("count", "sum", "sumsq", "min", "max", "mean", "stddev", "nonfinite")]


class UGridParams(ctypes.Structure):
    _fields_ = [("rows", c_size_t), ("cols", c_size_t)] + [(k, c_double) for k in (
        "mass", "radius", "spacing", "stiffness", "rest_length", "damping", "min_distance", "repulsion_stiffness",
        "h_attach0", "h_attach1", "v_attach0", "v_attach1")] + [("row_begin", c_size_t), ("rows_local", c_size_t), ("topology", c_int)] + [
        (k, c_double) for k in ("diag_rest_length", "d_attach0", "d_attach1", "a_attach0", "a_attach1")]


class Grid2DParams(ctypes.Structure):
    _fields_ = [("rows", c_size_t), ("cols", c_size_t), ("row_begin", c_size_t), ("rows_local", c_size_t),
                ("topology", c_int), ("mass", c_double), ("spacing", c_double), ("stiffness", c_double),
                ("damping", c_double)]


# every symbol include/sopot_b200.h declares: name -> (restype, argtypes)
_ctx = c_void_p
SYMBOLS = {
    "sopot_b200_abi_version": (c_int, []),
    "sopot_b200_nccl_unique_id": (c_int, [c_void_p]),
    "sopot_b200_init": (c_int, [POINTER(_ctx), c_int, c_void_p, c_int, c_int]),
    "sopot_b200_destroy": (None, [_ctx]),
    "sopot_b200_last_error": (c_char_p, [_ctx]),
    "sopot_b200_sync": (c_int, [_ctx]),
    "sopot_b200_get_stream": (c_int, [_ctx, POINTER(c_void_p)]),
    "sopot_b200_note_launches": (c_int, [_ctx, ctypes.c_uint64]),
    "sopot_b200_set_stream": (c_int, [_ctx, c_void_p]),
    "sopot_b200_device_info": (c_int, [_ctx, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                       POINTER(c_size_t), c_char_p, c_size_t]),
    "sopot_b200_malloc": (c_int, [_ctx, c_size_t, POINTER(c_void_p)]),
    "sopot_b200_free": (c_int, [_ctx, c_void_p]),
    "sopot_b200_host_alloc": (c_int, [_ctx, c_size_t, POINTER(c_void_p)]),
    "sopot_b200_host_free": (c_int, [_ctx, c_void_p]),
    "sopot_b200_memcpy_h2d": (c_int, [_ctx, c_void_p, c_void_p, c_size_t]),
    "sopot_b200_memcpy_d2h": (c_int, [_ctx, c_void_p, c_void_p, c_size_t]),
    "sopot_b200_memcpy_d2d": (c_int, [_ctx, c_void_p, c_void_p, c_size_t]),
    "sopot_b200_memset": (c_int, [_ctx, c_void_p, c_int, c_size_t]),
    "sopot_b200_event_record": (c_int, [_ctx, c_int]),
    "sopot_b200_event_elapsed_ms": (c_int, [_ctx, c_int, c_int, POINTER(c_float)]),
    "sopot_b200_launch_count": (c_uint64, [_ctx]),
    "sopot_b200_rk4_num_steps": (c_size_t, [c_double, c_double, c_double]),
    "sopot_b200_ho_rk4": (c_int, [_ctx, POINTER(HoParams), c_size_t, c_double, c_double, c_size_t,
                                  POINTER(RK4Opts), c_void_p, c_void_p, c_void_p]),
    "sopot_b200_ho_rhs": (c_int, [_ctx, POINTER(HoParams), c_size_t, POINTER(RK4Opts), c_void_p, c_void_p]),
    "sopot_b200_cartpole_rhs": (c_int, [_ctx, POINTER(CartpoleParams), c_size_t, POINTER(RK4Opts), c_void_p, c_void_p]),
    "sopot_b200_lqr4_gain": (c_int, [_ctx, c_void_p, c_void_p, c_size_t, c_void_p, c_double, c_double, c_size_t, c_double,
                                     POINTER(RK4Opts), c_void_p, c_void_p, c_void_p, c_void_p]),
    "sopot_b200_ugrid_initial_state": (c_int, [_ctx, POINTER(UGridParams), c_uint32, c_void_p]),
    "sopot_b200_ugrid_rhs": (c_int, [_ctx, POINTER(UGridParams), POINTER(RK4Opts), c_void_p, c_void_p]),
    "sopot_b200_ugrid_rk4": (c_int, [_ctx, POINTER(UGridParams), c_double, c_size_t, POINTER(RK4Opts), c_void_p]),
    "sopot_b200_ugrid_integrate": (c_int, [_ctx, POINTER(UGridParams), c_int, c_double, c_size_t, POINTER(RK4Opts), c_void_p]),
    "sopot_b200_lqr_gain": (c_int, [_ctx, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_double, c_double, c_size_t, c_double,
                                    POINTER(RK4Opts), c_void_p, c_void_p, c_void_p, c_void_p]),
    "sopot_b200_cartpole_feedback_rk4": (c_int, [_ctx, POINTER(CartpoleParams), POINTER(Feedback), c_size_t, c_double,
                                                 c_double, c_size_t, POINTER(RK4Opts), c_void_p, c_void_p, c_void_p,
                                                 c_void_p]),
    "sopot_b200_dpend_rk4": (c_int, [_ctx, POINTER(DpendParams), c_size_t, c_double, c_double, c_size_t,
                                     POINTER(RK4Opts), c_void_p, c_void_p, c_void_p]),
    "sopot_b200_dpend_rhs": (c_int, [_ctx, POINTER(DpendParams), c_size_t, POINTER(RK4Opts), c_void_p, c_void_p]),
    "sopot_b200_dpend_jacobian": (c_int, [_ctx, POINTER(DpendParams), c_size_t, POINTER(RK4Opts), c_void_p, c_void_p]),
    "sopot_b200_cartpole_rk4": (c_int, [_ctx, POINTER(CartpoleParams), c_size_t, c_double, c_double, c_size_t,
                                        POINTER(RK4Opts), c_void_p, c_void_p, c_void_p]),
    "sopot_b200_cartpole_linearize": (c_int, [_ctx, POINTER(CartpolePlant), c_void_p, c_void_p, c_size_t,
                                              POINTER(RK4Opts), c_void_p, c_void_p, c_void_p]),
    "sopot_b200_rocket_rk4": (c_int, [_ctx, POINTER(RocketParams), c_size_t, c_double, c_size_t,
                                      POINTER(RK4Opts), c_void_p, c_void_p, c_void_p, c_void_p]),
    "sopot_b200_rocket_rhs": (c_int, [_ctx, POINTER(RocketParams), c_size_t, POINTER(RK4Opts), c_void_p, c_void_p]),
    "sopot_b200_rocket_state_functions": (c_int, [_ctx, POINTER(RocketParams), c_size_t, POINTER(RK4Opts), c_void_p, c_void_p]),
    "sopot_b200_dispersion_stats": (c_int, [_ctx, c_void_p, c_size_t, c_size_t, c_uint32, POINTER(Dispersion)]),
    "sopot_b200_grid2d_initial_state": (c_int, [_ctx, POINTER(Grid2DParams), c_uint32, c_void_p]),
    "sopot_b200_grid2d_rk4": (c_int, [_ctx, POINTER(Grid2DParams), c_double, c_size_t, POINTER(RK4Opts), c_void_p]),
    "sopot_b200_grid2d_rhs": (c_int, [_ctx, POINTER(Grid2DParams), POINTER(RK4Opts), c_void_p, c_void_p]),
    "sopot_b200_measure_fp64_peak": (c_int, [_ctx, POINTER(c_double)]),
    "sopot_b200_measure_hbm_peak": (c_int, [_ctx, c_size_t, POINTER(c_double)]),
}

_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen libsopot_b200.so and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SopotB200Error(f"{LIB_PATH} is missing: build it with `python -m sopot_b200.build` "
                             "(the engine has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype, fn.argtypes = res, args
    if lib.sopot_b200_abi_version() != ABI_VERSION:
        raise SopotB200Error("ABI version mismatch")
    _lib = lib
    return lib


class DeviceArray:
    """float64 (or int32) array in device memory owned by an Engine."""

    def __init__(self, eng: "Engine", shape, dtype=np.float64):
        self.eng, self.shape, self.dtype = eng, tuple(np.atleast_1d(shape)), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = c_void_p()
        eng._ck(eng.lib.sopot_b200_malloc(eng.ctx, self.nbytes, byref(p)))
        self.ptr = p.value

    def upload(self, a) -> "DeviceArray":
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.nbytes == self.nbytes, (a.shape, self.shape)
        self._keep = a
        self.eng._ck(self.eng.lib.sopot_b200_memcpy_h2d(self.eng.ctx, self.ptr, a.ctypes.data, self.nbytes))
        self.eng.sync()
        self._keep = None
        return self

    def download(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        self.eng._ck(self.eng.lib.sopot_b200_memcpy_d2h(self.eng.ctx, out.ctypes.data, self.ptr, self.nbytes))
        self.eng.sync()
        return out

    def free(self):
        if self.ptr:
            self.eng.lib.sopot_b200_free(self.eng.ctx, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            if self.eng.ctx:
                self.free()
        except Exception:
            pass


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, DeviceArray):
        return a.ptr
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    raise TypeError(type(a))


class Engine:
    def __init__(self, device: int = 0, nccl_id: bytes | None = None, rank: int = 0, nranks: int = 1):
        self.lib = load_library()
        ctx = c_void_p()
        rc = self.lib.sopot_b200_init(byref(ctx), device, nccl_id, rank, nranks)
        if rc != 0:
            raise SopotB200Error(f"sopot_b200_init failed ({rc}): "
                                 f"{self.lib.sopot_b200_last_error(None).decode()}")
        self.ctx, self.rank, self.nranks = ctx, rank, nranks

    # ---- plumbing -------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != 0:
            raise SopotB200Error(f"[{rc}] {self.lib.sopot_b200_last_error(self.ctx).decode()}")

    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = load_library()
        buf = ctypes.create_string_buffer(128)
        rc = lib.sopot_b200_nccl_unique_id(buf)
        if rc != 0:
            raise SopotB200Error(lib.sopot_b200_last_error(None).decode())
        return buf.raw

    def close(self):
        if self.ctx:
            self.lib.sopot_b200_destroy(self.ctx)
            self.ctx = None

    def sync(self):
        self._ck(self.lib.sopot_b200_sync(self.ctx))

    def set_stream(self, cuda_stream: int | None):
        self._ck(self.lib.sopot_b200_set_stream(self.ctx, cuda_stream))

    def device_info(self) -> dict:
        sm, ma, mi, mem = c_int(), c_int(), c_int(), c_size_t()
        name = ctypes.create_string_buffer(256)
        self._ck(self.lib.sopot_b200_device_info(self.ctx, byref(sm), byref(ma), byref(mi), byref(mem), name, 256))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), mem_bytes=mem.value, name=name.value.decode())

    def empty(self, shape, dtype=np.float64) -> DeviceArray:
        return DeviceArray(self, shape, dtype)

    def to_device(self, a, dtype=np.float64) -> DeviceArray:
        a = np.ascontiguousarray(a, dtype=dtype)
        return DeviceArray(self, a.shape, dtype).upload(a)

    def pinned(self, shape, dtype=np.float64) -> np.ndarray:
        """numpy view of page-locked host memory (freed with the process)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = c_void_p()
        self._ck(self.lib.sopot_b200_host_alloc(self.ctx, max(n, 1), byref(p)))
        buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def event_record(self, slot: int):
        self._ck(self.lib.sopot_b200_event_record(self.ctx, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = c_float()
        self._ck(self.lib.sopot_b200_event_elapsed_ms(self.ctx, a, b, byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        return int(self.lib.sopot_b200_launch_count(self.ctx))

    def rk4_num_steps(self, t0, t1, dt) -> int:
        return int(self.lib.sopot_b200_rk4_num_steps(t0, t1, dt))

    # ---- helpers --------------------------------------------------------------------------
    def _prep(self, arrays: dict, n: int, mem: int):
        """Per-instance parameter dict -> (pointers, broadcast mask, keep-alive list)."""
        ptrs, mask, keep = {}, 0, []
        for j, (k, v) in enumerate(arrays.items()):
            if isinstance(v, DeviceArray):
                assert mem == MEM_DEVICE
                if int(np.prod(v.shape)) != 1 or n == 1:
                    ptrs[k] = v.ptr
                    continue
                v = v.download()                      # broadcast scalars are host values by contract
            a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
            if a.ndim == 0 or (a.size == 1 and n != 1):
                a = a.reshape(1)
                mask |= 1 << j
                ptrs[k] = a.ctypes.data               # host scalar, read by the entry point at call time
                keep.append(a)
                continue
            assert a.size == n, (k, a.size, n)
            if mem == MEM_DEVICE:
                a = self.to_device(a)
                ptrs[k] = a.ptr
            else:
                ptrs[k] = a.ctypes.data
            keep.append(a)
        return ptrs, mask, keep

    def _outs(self, dim, n, n_steps, store_every, mem, want_status, state_out=None, traj_out=None):
        snaps = n_steps // store_every if store_every else 0
        if mem == MEM_HOST:
            state = np.empty((dim, n)) if state_out is None else state_out
            traj = (np.empty((snaps, dim, n)) if traj_out is None else traj_out) if snaps else None
            status = np.zeros(n, dtype=np.int32) if want_status else None
        else:
            state = self.empty((dim, n)) if state_out is None else state_out
            traj = (self.empty((snaps, dim, n)) if traj_out is None else traj_out) if snaps else None
            status = self.empty((n,), np.int32) if want_status else None
        return state, traj, status

    def _ensemble(self, fn, pstruct, arrays, dim, n, t0, dt, n_steps, mem, fp_mode, store_every,
                  want_status, state_out, traj_out, with_t0=True):
        ptrs, mask, keep = self._prep(arrays, n, mem)
        p = pstruct(**ptrs)
        opts = RK4Opts(mem, fp_mode, store_every, mask, 0)
        state, traj, status = self._outs(dim, n, n_steps, store_every, mem, want_status, state_out, traj_out)
        args = [self.ctx, byref(p), n] + ([t0] if with_t0 else []) + [dt, n_steps, byref(opts), _ptr(state), _ptr(traj), _ptr(status)]
        self._ck(fn(*args))
        self._keep = keep
        return state, traj, status

    # ---- entry points -----------------------------------------------------------------------
    def ho_rk4(self, m, k, c, x0, v0, n, t0, dt, n_steps, mem=MEM_HOST, fp_mode=FP_CONTRACT, store_every=0,
               want_status=True, state_out=None, traj_out=None, sync=True):
        r = self._ensemble(self.lib.sopot_b200_ho_rk4, HoParams, dict(m=m, k=k, c=c, x0=x0, v0=v0), 2, n, t0, dt,
                           n_steps, mem, fp_mode, store_every, want_status, state_out, traj_out)
        if sync:
            self.sync()
        return r

    def dpend_rk4(self, m1, m2, L1, L2, g, th1, th2, w1, w2, n, t0, dt, n_steps, mem=MEM_HOST,
                  fp_mode=FP_CONTRACT, store_every=0, want_status=True, state_out=None, traj_out=None, sync=True):
        r = self._ensemble(self.lib.sopot_b200_dpend_rk4, DpendParams,
                           dict(m1=m1, m2=m2, L1=L1, L2=L2, g=g, th1=th1, th2=th2, w1=w1, w2=w2), 4, n, t0, dt,
                           n_steps, mem, fp_mode, store_every, want_status, state_out, traj_out)
        if sync:
            self.sync()
        return r

    def dpend_rhs(self, m1, m2, L1, L2, g, state, fp_mode=FP_CONTRACT):
        state = np.ascontiguousarray(state, dtype=np.float64)
        n = state.shape[1]
        ptrs, mask, keep = self._prep(dict(m1=m1, m2=m2, L1=L1, L2=L2, g=g), n, MEM_HOST)
        p = DpendParams(**ptrs)
        opts = RK4Opts(MEM_HOST, fp_mode, 0, mask, 0)
        out = np.empty((4, n))
        self._ck(self.lib.sopot_b200_dpend_rhs(self.ctx, byref(p), n, byref(opts), _ptr(state), _ptr(out)))
        self.sync()
        return out

    def dpend_jacobian(self, m1, m2, L1, L2, g, state, fp_mode=FP_STRICT, mem=MEM_HOST, J=None, sync=True):
        """d f_i / d x_j of the double pendulum at every column of state[4][n] -> J[16][n] (row-major 4x4 per point)."""
        if mem == MEM_HOST:
            state = np.ascontiguousarray(state, dtype=np.float64)
        n = state.shape[1]
        ptrs, mask, keep = self._prep(dict(m1=m1, m2=m2, L1=L1, L2=L2, g=g), n, mem)
        p = DpendParams(**ptrs)
        opts = RK4Opts(mem, fp_mode, 0, mask, 0)
        if J is None:
            J = np.empty((16, n)) if mem == MEM_HOST else self.empty((16, n))
        self._ck(self.lib.sopot_b200_dpend_jacobian(self.ctx, byref(p), n, byref(opts), _ptr(state), _ptr(J)))
        self._keep = keep
        if sync:
            self.sync()
        return J

    def cartpole_rk4(self, mc, m, L, g, x, th, xd, w, force, n, t0, dt, n_steps, mem=MEM_HOST,
                     fp_mode=FP_CONTRACT, store_every=0, want_status=True, sync=True):
        r = self._ensemble(self.lib.sopot_b200_cartpole_rk4, CartpoleParams,
                           dict(mc=mc, m=m, L=L, g=g, x=x, th=th, xd=xd, w=w, force=force), 4, n, t0, dt, n_steps,
                           mem, fp_mode, store_every, want_status, None, None)
        if sync:
            self.sync()
        return r

    def cartpole_linearize(self, mc, m, L, g, x0, u0, P, mem=MEM_HOST, fp_mode=FP_CONTRACT, A=None, B=None,
                           want_status=True, sync=True):
        ptrs, mask, keep = self._prep(dict(mc=mc, m=m, L=L, g=g), P, mem)
        plant = CartpolePlant(**ptrs)
        opts = RK4Opts(mem, fp_mode, 0, mask, 0)
        if mem == MEM_HOST:
            x0 = np.ascontiguousarray(x0, dtype=np.float64); u0 = np.ascontiguousarray(u0, dtype=np.float64)
            A = np.empty((16, P)) if A is None else A
            B = np.empty((4, P)) if B is None else B
            status = np.zeros(P, dtype=np.int32) if want_status else None
        else:
            A = self.empty((16, P)) if A is None else A
            B = self.empty((4, P)) if B is None else B
            status = self.empty((P,), np.int32) if want_status else None
        self._ck(self.lib.sopot_b200_cartpole_linearize(self.ctx, byref(plant), _ptr(x0), _ptr(u0), P, byref(opts),
                                                        _ptr(A), _ptr(B), _ptr(status)))
        self._keep = keep
        if sync:
            self.sync()
        return A, B, status

    def lqr4_gain(self, A, B, P, Q, R, riccati_dt=0.01, max_iter=1000, tol=1e-8, mem=MEM_HOST, fp_mode=FP_CONTRACT,
                  K=None, want_riccati=False, want_iters=True, sync=True):
        """Batched LQR<4,1>::solve on linearizer output A [16][P], B [4][P] -> K [4][P], (Pric, iters, status)."""
        Q = np.ascontiguousarray(np.asarray(Q, dtype=np.float64).reshape(16))
        opts = RK4Opts(mem, fp_mode, 0, 0, 0)
        if mem == MEM_HOST:
            A = np.ascontiguousarray(A, dtype=np.float64); B = np.ascontiguousarray(B, dtype=np.float64)
            K = np.empty((4, P)) if K is None else K
            Pric = np.empty((16, P)) if want_riccati else None
            iters = np.zeros(P, dtype=np.int32) if want_iters else None
            status = np.zeros(P, dtype=np.int32)
        else:
            K = self.empty((4, P)) if K is None else K
            Pric = self.empty((16, P)) if want_riccati else None
            iters = self.empty((P,), np.int32) if want_iters else None
            status = self.empty((P,), np.int32)
        self._ck(self.lib.sopot_b200_lqr4_gain(self.ctx, _ptr(A), _ptr(B), P, Q.ctypes.data, R, riccati_dt, max_iter, tol,
                                               byref(opts), _ptr(K), _ptr(Pric), _ptr(iters), _ptr(status)))
        if sync:
            self.sync()
        return K, Pric, iters, status

    def lqr_gain(self, state_dim, A, B, P, Q, R, riccati_dt=0.01, max_iter=1000, tol=1e-8, mem=MEM_HOST, fp_mode=FP_CONTRACT,
                 want_riccati=False, sync=True):
        """Batched LQR<state_dim,1>::solve (state_dim 4, 6, 14): A [sd^2][P], B [sd][P] -> K [sd][P], Pric, iters, status."""
        Q = np.ascontiguousarray(np.asarray(Q, dtype=np.float64).reshape(state_dim * state_dim))
        opts = RK4Opts(mem, fp_mode, 0, 0, 0)
        if mem == MEM_HOST:
            A = np.ascontiguousarray(A, dtype=np.float64); B = np.ascontiguousarray(B, dtype=np.float64)
            K = np.empty((state_dim, P)); Pric = np.empty((state_dim * state_dim, P)) if want_riccati else None
            iters = np.zeros(P, dtype=np.int32); status = np.zeros(P, dtype=np.int32)
        else:
            K = self.empty((state_dim, P)); Pric = self.empty((state_dim * state_dim, P)) if want_riccati else None
            iters = self.empty((P,), np.int32); status = self.empty((P,), np.int32)
        self._ck(self.lib.sopot_b200_lqr_gain(self.ctx, state_dim, _ptr(A), _ptr(B), P, Q.ctypes.data, R, riccati_dt, max_iter, tol,
                                              byref(opts), _ptr(K), _ptr(Pric), _ptr(iters), _ptr(status)))
        if sync:
            self.sync()
        return K, Pric, iters, status

    def cartpole_feedback_rk4(self, mc, m, L, g, x, th, xd, w, K, n, t0, dt, n_steps, x_ref=None, u_min=0.0, u_max=0.0,
                              saturate=False, mem=MEM_HOST, fp_mode=FP_CONTRACT, store_every=0, want_status=True,
                              want_stats=True, sync=True, state_out=None):
        """Closed-loop cart-pole ensemble: u = clamp(-K (x - x_ref)) held over each RK4 step."""
        ptrs, mask, keep = self._prep(dict(mc=mc, m=m, L=L, g=g, x=x, th=th, xd=xd, w=w), n, mem)
        params = CartpoleParams(force=None, **ptrs)
        if mem == MEM_HOST:
            K = np.ascontiguousarray(K, dtype=np.float64)
            x_ref = None if x_ref is None else np.ascontiguousarray(x_ref, dtype=np.float64)
        fb = Feedback(_ptr(K), _ptr(x_ref), u_min, u_max, int(bool(saturate)))
        opts = RK4Opts(mem, fp_mode, store_every, mask, 0)
        state, traj, status = self._outs(4, n, n_steps, store_every, mem, want_status, state_out, None)
        stats = None
        if want_stats:
            stats = np.empty((2, n)) if mem == MEM_HOST else self.empty((2, n))
        self._ck(self.lib.sopot_b200_cartpole_feedback_rk4(self.ctx, byref(params), byref(fb), n, t0, dt, n_steps,
                                                           byref(opts), _ptr(state), _ptr(traj), _ptr(stats), _ptr(status)))
        self._keep = keep
        if sync:
            self.sync()
        return state, stats, traj, status

    @staticmethod
    def aero_tables(tables: dict, engine_on: bool = True):
        """{name: (row_vals, col_vals, data[rows][cols])} for names in AERO_TABLE_NAMES -> (AeroTables, keep-alive)."""
        a, keep = AeroTables(), []
        a.engine_on = int(bool(engine_on))
        for name, (rv, cv, data) in tables.items():
            assert name in AERO_TABLE_NAMES, name
            rv, cv, data = (np.ascontiguousarray(x, dtype=np.float64) for x in (rv, cv, data))
            assert data.shape == (rv.size, cv.size)
            keep += [rv, cv, data]
            setattr(a, name, Table2D(rv.size, cv.size, rv.ctypes.data, cv.ctypes.data, data.ctypes.data))
        return a, keep

    def _rocket_params(self, elev, az, diam, g0, engine, mass_tab, ix_tab, iyz_tab, n, mem, aero=None):
        ptrs, mask, keep = self._prep(dict(elevation_deg=elev, azimuth_deg=az, diameter=diam, gravity=g0), n, mem)
        tabs = {}
        for name, t, cols in (("engine", engine, 8), ("mass", mass_tab, 2), ("ix", ix_tab, 2), ("iyz", iyz_tab, 2)):
            if t is None:
                tabs[name + "_rows"], tabs[name] = 0, None
            else:
                t = np.ascontiguousarray(t, dtype=np.float64)
                assert t.ndim == 2 and t.shape[1] == cols
                keep.append(t)
                tabs[name + "_rows"], tabs[name] = t.shape[0], t.ctypes.data
        p = RocketParams(**ptrs, **tabs)
        if aero is not None:
            a, k2 = aero if isinstance(aero, tuple) else self.aero_tables(aero)
            keep += k2 + [a]
            p.aero = ctypes.pointer(a)
        return p, mask, keep

    def rocket_rk4(self, elev, az, diam, g0, engine, mass_tab, ix_tab, iyz_tab, n, dt, n_steps, mem=MEM_HOST,
                   fp_mode=FP_CONTRACT, store_every=0, want_status=True, want_stats=True, state_out=None,
                   stats_out=None, sync=True, aero=None):
        p, mask, keep = self._rocket_params(elev, az, diam, g0, engine, mass_tab, ix_tab, iyz_tab, n, mem, aero)
        opts = RK4Opts(mem, fp_mode, store_every, mask, 0)
        state, traj, status = self._outs(14, n, n_steps, store_every, mem, want_status, state_out, None)
        stats = None
        if want_stats:
            stats = stats_out if stats_out is not None else (np.empty((8, n)) if mem == MEM_HOST else self.empty((8, n)))
        self._ck(self.lib.sopot_b200_rocket_rk4(self.ctx, byref(p), n, dt, n_steps, byref(opts), _ptr(state),
                                                _ptr(traj), _ptr(stats), _ptr(status)))
        self._keep = keep
        if sync:
            self.sync()
        return state, stats, traj, status

    def rocket_rhs(self, elev, az, diam, g0, engine, mass_tab, state, fp_mode=FP_CONTRACT, ix_tab=None, iyz_tab=None, aero=None):
        state = np.ascontiguousarray(state, dtype=np.float64)
        n = state.shape[1]
        p, mask, keep = self._rocket_params(elev, az, diam, g0, engine, mass_tab, ix_tab, iyz_tab, n, MEM_HOST, aero)
        opts = RK4Opts(MEM_HOST, fp_mode, 0, mask, 0)
        out = np.empty((14, n))
        self._ck(self.lib.sopot_b200_rocket_rhs(self.ctx, byref(p), n, byref(opts), _ptr(state), _ptr(out)))
        self.sync()
        return out

    ROCKET_PROBE = dict(mass=0, inertia=slice(1, 4), total_force_enu=slice(4, 7), total_torque_body=slice(7, 10), gravity=10,
                        aero_force_body=slice(11, 14), aero_force_enu=slice(14, 17), aero_moment_body=slice(17, 20), density=20,
                        pressure=21, temperature=22, speed_of_sound=23, mass_flow_rate=24, thrust_force_body=slice(25, 28))

    def rocket_state_functions(self, elev, az, diam, g0, engine, mass_tab, state, ix_tab=None, iyz_tab=None, aero=None):
        """The reference's state functions at state[14][n] -> [28][n] (rows: Engine.ROCKET_PROBE)."""
        state = np.ascontiguousarray(state, dtype=np.float64)
        n = state.shape[1]
        p, mask, keep = self._rocket_params(elev, az, diam, g0, engine, mass_tab, ix_tab, iyz_tab, n, MEM_HOST, aero)
        opts = RK4Opts(MEM_HOST, FP_STRICT, 0, mask, 0)
        out = np.empty((28, n))
        self._ck(self.lib.sopot_b200_rocket_state_functions(self.ctx, byref(p), n, byref(opts), _ptr(state), _ptr(out)))
        self.sync()
        return out

    def dispersion_stats(self, values, rows=None, n_local=None):
        if isinstance(values, DeviceArray):
            mem = MEM_DEVICE
            rows, n_local = values.shape if rows is None else (rows, n_local)
        else:
            values = np.ascontiguousarray(values, dtype=np.float64)
            mem = MEM_HOST
            rows, n_local = values.shape
        out = (Dispersion * rows)()
        self._ck(self.lib.sopot_b200_dispersion_stats(self.ctx, _ptr(values), rows, n_local, mem, out))
        return [{k: getattr(o, k) for k, _ in Dispersion._fields_} for o in out]

    # ---- unified 6-state grid ----
    @staticmethod
    def ugrid_params(rows, cols, mass, radius, spacing, k, L0, c, min_distance=0.0, krep=-1.0,
                     attach=(0.0, 3.14159265358979323846, 3.14159265358979323846 / 2.0, -3.14159265358979323846 / 2.0),
                     row_begin=0, rows_local=0, triangle=False, diag_rest_length=None,
                     diag_attach=(3.14159265358979323846 / 4.0, -3.0 * 3.14159265358979323846 / 4.0,
                                  3.0 * 3.14159265358979323846 / 4.0, -3.14159265358979323846 / 4.0)):
        """rows_local = 0: the whole grid on this rank; else this rank's slab of rows (multi-GPU, like grid2d).
        triangle: add the two diagonal springs of every cell (the simulator's "triangle" grid type); their rest length
        defaults to sqrt(2) * spacing as in the reference wrapper."""
        Ld = float(np.sqrt(2.0) * spacing) if diag_rest_length is None else diag_rest_length
        return UGridParams(rows, cols, mass, radius, spacing, k, L0, c, min_distance, krep, *attach, row_begin, rows_local,
                           1 if triangle else 0, Ld, *diag_attach)

    def ugrid_initial_state(self, g, mem=MEM_HOST):
        nm = (g.rows_local or g.rows) * g.cols
        out = np.empty((nm, 6)) if mem == MEM_HOST else self.empty((nm, 6))
        self._ck(self.lib.sopot_b200_ugrid_initial_state(self.ctx, byref(g), mem, _ptr(out)))
        self.sync()
        return out

    def ugrid_rhs(self, g, state, fp_mode=FP_STRICT):
        mem = MEM_DEVICE if isinstance(state, DeviceArray) else MEM_HOST
        opts = RK4Opts(mem, fp_mode, 0, 0, 0)
        out = np.empty_like(state) if mem == MEM_HOST else self.empty(state.shape)
        self._ck(self.lib.sopot_b200_ugrid_rhs(self.ctx, byref(g), byref(opts), _ptr(state), _ptr(out)))
        self.sync()
        return out

    def ugrid_rk4(self, g, dt, n_steps, state, fp_mode=FP_STRICT, sync=True, integrator=0, stage_pairs=False, kernel=None):
        """integrator: 0 RK4, 1 symplectic Euler, 2 velocity Verlet (the reference Grid2DSimulator's three).
        stage_pairs: RK4 with two stages per kernel instead of the default one kernel per stage (single rank, opt-in).
        kernel="march" / "gather": force the stage kernel of the quad topology (default: by mode, integrator and size)."""
        mem = MEM_DEVICE if isinstance(state, DeviceArray) else MEM_HOST
        flags = (FLAG_UGRID_STAGE_PAIRS if stage_pairs else 0) | {None: 0, "march": FLAG_UGRID_MARCH, "gather": FLAG_UGRID_GATHER}[kernel]
        opts = RK4Opts(mem, fp_mode, 0, 0, flags)
        self._ck(self.lib.sopot_b200_ugrid_integrate(self.ctx, byref(g), integrator, dt, n_steps, byref(opts), _ptr(state)))
        if sync:
            self.sync()
        return state

    def grid2d_params(self, rows, cols, topo, mass, spacing, k, c, row_begin=0, rows_local=None):
        return Grid2DParams(rows, cols, row_begin, rows if rows_local is None else rows_local, topo, mass,
                            spacing, k, c)

    def grid2d_initial_state(self, g: Grid2DParams, mem=MEM_HOST):
        st = np.empty((g.rows_local * g.cols, 4)) if mem == MEM_HOST else self.empty((g.rows_local * g.cols, 4))
        self._ck(self.lib.sopot_b200_grid2d_initial_state(self.ctx, byref(g), mem, _ptr(st)))
        self.sync()
        return st

    def grid2d_rk4(self, g: Grid2DParams, dt, n_steps, state, fp_mode=FP_STRICT, sync=True, per_stage=False, kernel=None):
        """In place on `state` (numpy host array or DeviceArray).  per_stage=True: four stage kernels with a 1-row
        halo exchange each instead of the fused step kernel (same values).  kernel="march" / "tile": force the fused step's
        kernel for the quad topology (default: chosen by size)."""
        mem = MEM_DEVICE if isinstance(state, DeviceArray) else MEM_HOST
        flags = (FLAG_GRID_PER_STAGE if per_stage else 0) | {None: 0, "march": FLAG_GRID_MARCH, "tile": FLAG_GRID_TILE}[kernel]
        opts = RK4Opts(mem, fp_mode, 0, 0, flags)
        self._ck(self.lib.sopot_b200_grid2d_rk4(self.ctx, byref(g), dt, n_steps, byref(opts), _ptr(state)))
        if sync:
            self.sync()
        return state

    def grid2d_rhs(self, g: Grid2DParams, state, fp_mode=FP_STRICT):
        mem = MEM_DEVICE if isinstance(state, DeviceArray) else MEM_HOST
        opts = RK4Opts(mem, fp_mode, 0, 0, 0)
        out = np.empty_like(state) if mem == MEM_HOST else self.empty(state.shape)
        self._ck(self.lib.sopot_b200_grid2d_rhs(self.ctx, byref(g), byref(opts), _ptr(state), _ptr(out)))
        self.sync()
        return out

    def measure_fp64_peak(self) -> float:
        v = c_double()
        self._ck(self.lib.sopot_b200_measure_fp64_peak(self.ctx, byref(v)))
        return v.value

    def measure_hbm_peak(self, nbytes: int = 1 << 31) -> float:
        v = c_double()
        self._ck(self.lib.sopot_b200_measure_hbm_peak(self.ctx, nbytes, byref(v)))
        return v.value
