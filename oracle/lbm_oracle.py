"""ctypes front-end of oracle/liblbm_oracle.so (TEST INFRASTRUCTURE ONLY — see lbm_oracle.c).

`Oracle` mirrors the method names of the reference's `pub struct LBM` (lbm-wgpu/src/lbm.rs:32-98,
methods :726, :1065, :1076, :1090, :1118, :1127, :1337-1365, :1502) so parity tests read like the
product's API.  Parity status: pinned to outputs of the reference's shipped binary for the host-side rows and to
the executed WGSL text for the device passes; unpinned against a real wgpu run (see lbm_oracle.c).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# LBM_ORACLE_CONTRACT=1 selects the contraction twin (see the header of lbm_oracle.c); the parity tests then pair it
# with lbm_b200/libblbm_contract.so (BLBM_LIBRARY)
CONTRACT = os.environ.get("LBM_ORACLE_CONTRACT", "0") not in ("", "0")
_LIB_PATH = os.path.join(_HERE, "liblbm_oracle_contract.so" if CONTRACT else "liblbm_oracle.so")

POP_NAMES = ("nw", "n", "ne", "w", "rest", "e", "sw", "s", "se")  # lbm.rs:632-640
CURL, UX, UY, RHO, SPEED = range(5)  # lbm.rs:10-16


def build(force=False):
    """Compile the C oracle in place (gcc, a second or two)."""
    src = os.path.join(_HERE, "lbm_oracle.c")
    twin = os.path.join(_HERE, "liblbm_oracle_contract.so")
    if force or any(not os.path.exists(p) or os.path.getmtime(p) < os.path.getmtime(src) for p in (_LIB_PATH, twin)):
        subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_libs = {}


def lib(contract=None):
    """the oracle library: the default build, or (contract=True / LBM_ORACLE_CONTRACT=1) the contraction twin"""
    contract = CONTRACT if contract is None else bool(contract)
    _lib = _libs.get(contract)
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liblbm_oracle_contract.so" if contract else "liblbm_oracle.so"))
        P = C.c_void_p
        L.lbm_oracle_create.restype = P
        L.lbm_oracle_create.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float]
        L.lbm_oracle_destroy.argtypes = [P]
        for name in ("collide", "stream", "step", "reset_barrier", "reset_to_equilibrium"):
            getattr(L, "lbm_oracle_" + name).argtypes = [P]
            getattr(L, "lbm_oracle_" + name).restype = None
        L.lbm_oracle_iterate.argtypes = [P, C.c_uint32]
        L.lbm_oracle_summary.argtypes = [P, C.c_int]
        L.lbm_oracle_set_summary.argtypes = [P, C.c_int]
        L.lbm_oracle_draw_points.argtypes = [P, P, C.c_size_t]
        L.lbm_oracle_set_omega.argtypes = [P, C.c_float]
        L.lbm_oracle_custom_speed.argtypes = [P, C.c_float]
        L.lbm_oracle_single_cell.argtypes = [P, C.c_uint32]
        L.lbm_oracle_cell_class.argtypes = [P, P]
        L.lbm_oracle_set_equil.argtypes = [C.c_float, C.c_float, C.c_float, P]
        L.lbm_oracle_color_map.argtypes = [P, C.c_int, P]
        L.lbm_oracle_population.restype = P
        L.lbm_oracle_population.argtypes = [P, C.c_int, C.c_int]
        for name in ("barrier", "mx", "my", "rho", "output"):
            getattr(L, "lbm_oracle_" + name).restype = P
            getattr(L, "lbm_oracle_" + name).argtypes = [P]
        L.lbm_oracle_compute_num.restype = C.c_uint64
        L.lbm_oracle_compute_num.argtypes = [P]
        L.lbm_oracle_threads.restype = C.c_int
        L.lbm_oracle_set_threads.argtypes = [C.c_int]
        L.lbm_oracle_contract.restype = C.c_int
        L.lbm_oracle_set_oob_clamp.argtypes = [P, C.c_int]
        assert bool(L.lbm_oracle_contract()) == contract
        _libs[contract] = _lib = L
    return _lib


def set_equil(ux, uy, rho):
    out = np.zeros(9, np.float32)
    lib().lbm_oracle_set_equil(ux, uy, rho, out.ctypes.data)
    return out


def threads():
    return lib().lbm_oracle_threads()


def use_all_cores():
    """OpenMP threads = the cores this process may run on (torchrun sets OMP_NUM_THREADS=1 for every rank; the
    timed CPU arm runs on rank 0 alone).  Returns the thread count in effect."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().lbm_oracle_set_threads(n)
    return threads()


class Oracle:
    def __init__(self, omega, x, y, inflow_ux=0.1, contract=None, oob="zero"):
        """oob: what an out-of-range read returns - "zero" (the defined semantics) or "clamp"
        (the last element: the other behaviour WebGPU allows; sensitivity studies only, no CUDA counterpart)"""
        self.w, self.h = int(x), int(y)
        self._L = lib(contract)
        self._h = self._L.lbm_oracle_create(self.w, self.h, float(omega), float(inflow_ux))
        if not self._h:
            raise MemoryError("lbm_oracle_create failed")
        if oob not in ("zero", "clamp"):
            raise ValueError(oob)
        self._L.lbm_oracle_set_oob_clamp(self._h, 1 if oob == "clamp" else 0)

    def close(self):
        if self._h:
            self._L.lbm_oracle_destroy(self._h)
            self._h = None

    __del__ = close

    def _view(self, ptr, dtype):
        n = self.w * self.h
        buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(self.h, self.w)

    # --- the reference's API surface ---
    def iterate(self, n):
        self._L.lbm_oracle_iterate(self._h, int(n))

    def collide(self):
        self._L.lbm_oracle_collide(self._h)

    def stream(self):
        self._L.lbm_oracle_stream(self._h)

    def step(self):
        self._L.lbm_oracle_step(self._h)

    def set_summary(self, stat):
        self._L.lbm_oracle_set_summary(self._h, int(stat))

    def compute_summary(self, stat):
        self._L.lbm_oracle_summary(self._h, int(stat))

    def draw_points(self, pairs):
        a = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1)
        assert a.size % 2 == 0
        self._L.lbm_oracle_draw_points(self._h, a.ctypes.data, a.size // 2)

    def reset_barrier(self):
        self._L.lbm_oracle_reset_barrier(self._h)

    def update_omega_buffer(self, omega):
        self._L.lbm_oracle_set_omega(self._h, float(omega))

    def reset_to_equilibrium(self):
        self._L.lbm_oracle_reset_to_equilibrium(self._h)

    def custom_speed(self, ux):
        self._L.lbm_oracle_custom_speed(self._h, float(ux))

    def single_cell(self, index):
        self._L.lbm_oracle_single_cell(self._h, int(index))

    def get_compute_num(self):
        return int(self._L.lbm_oracle_compute_num(self._h))

    # --- read-back (views into the oracle's memory; copy if you need to keep them) ---
    def population(self, buffer, k):
        """data_buffers[buffer][k]; buffer -1 = the live one (compute_step % 2).  k=4 (rest) always
        reads buffer 0, the only rest array the reference binds (lbm.rs:775-778)."""
        if buffer < 0:
            buffer = self.get_compute_num() % 2
        if k == 4:
            buffer = 0
        return self._view(self._L.lbm_oracle_population(self._h, buffer, k), np.float32)

    def barrier(self):
        return self._view(self._L.lbm_oracle_barrier(self._h), np.uint32)

    def moments(self):
        return (self._view(self._L.lbm_oracle_mx(self._h), np.float32),
                self._view(self._L.lbm_oracle_my(self._h), np.float32),
                self._view(self._L.lbm_oracle_rho(self._h), np.float32))

    def output(self):
        return self._view(self._L.lbm_oracle_output(self._h), np.float32)

    def color_map(self, cmap):
        """colours of the current output field, (h, w, 3) fp32; Inferno 0, Viridis 1, Jet 2 (lbm.rs:18-24)"""
        out = np.zeros((self.h, self.w, 3), np.float32)
        self._L.lbm_oracle_color_map(self._h, int(cmap), out.ctypes.data)
        return out

    def cell_class(self):
        out = np.zeros((self.h, self.w), np.uint16)
        self._L.lbm_oracle_cell_class(self._h, out.ctypes.data)
        return out
