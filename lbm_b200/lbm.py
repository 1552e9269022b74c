"""Host-side mirror of the reference's `pub struct LBM` over the C ABI (include/blbm.h).

Method names, argument meaning and call order follow lbm-wgpu/src/lbm.rs so that code (and tests)
written against the reference reads the same here:

    reference (lbm.rs)                          here
    LBM::new(&driver, omega, x, y)     :726     LBM(omega, x, y)
    iterate(&driver, n)                :1065    iterate(n)
    collide(&driver) / stream(&driver) :1118    collide() / stream()
    rerender(&driver)                  :1104    rerender()
    set_summary(stat)                  :1061    set_summary(stat)
    draw_shape(&driver, &shape)        :1337    draw_shape(shape) / draw_points(pairs)
    reset_barrier(&driver)             :1362    reset_barrier()
    update_omega_buffer(&driver, w)    :1358    update_omega_buffer(w)
    reset_to_equilibrium(&driver)      :1076    reset_to_equilibrium()
    custom_speed(&driver, ux)          :1090    custom_speed(ux)
    single_cell(&driver, index)        :1502    single_cell(index)
    get_compute_num / get_frame_num    :1166    get_compute_num() / get_frame_num()

`&Driver` disappears (the handle owns the CUDA device and stream).  Where the reference panics, these
raise BlbmError.  Read-back methods are new: the reference never reads its buffers back.
"""
import ctypes as C
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
POP_NAMES = ("nw", "n", "ne", "w", "rest", "e", "sw", "s", "se")  # lbm.rs:632-640
PEER_HANDLE_BYTES = 512


class SummaryStat(enum.IntEnum):  # lbm.rs:10-16
    Curl = 0
    Ux = 1
    Uy = 2
    Rho = 3
    Speed = 4


class ColorMap(enum.IntEnum):  # lbm.rs:18-24
    Inferno = 0
    Viridis = 1
    Jet = 2


class Kernel(enum.IntEnum):
    Auto = 0
    Scalar = 1
    Vec4 = 2


class BlbmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"blbm error {code}: {msg}")
        self.code = code


def omega_from_viscosity(nu):
    """lib.rs:183: omega = 1 / (3 nu + 0.5), evaluated in f32 like the reference."""
    f = np.float32
    return float(f(1.0) / (f(3.0) * f(nu) + f(0.5)))


def library_path():
    return os.environ.get("BLBM_LIBRARY", os.path.join(_HERE, "libblbm.so"))


_lib = None

# name -> (restype, argtypes); the single source of truth for the ctypes prototypes.  tests/test_abi.py
# checks this table against include/blbm.h.
_P, _U32, _U64, _I, _F, _SZ = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float, C.c_size_t
PROTOTYPES = {
    "blbm_last_error": (C.c_char_p, []),
    "blbm_abi_version": (_I, []),
    "blbm_device_count": (_I, []),
    "blbm_create": (_I, [_U32, _U32, _F, _F, _I, C.POINTER(_P)]),
    "blbm_create_slab": (_I, [_U32, _U64, _U64, _U64, _F, _F, _I, C.POINTER(_P)]),
    "blbm_create_group": (_I, [_U32, _U64, _F, _F, C.POINTER(_I), _I, C.POINTER(_P)]),
    "blbm_group_size": (_I, [_P]),
    "blbm_group_slab": (_I, [_P, _I, C.POINTER(_P)]),
    "blbm_destroy": (_I, [_P]),
    "blbm_iterate": (_I, [_P, _U32]),
    "blbm_advance": (_I, [_P, _U32]),
    "blbm_iterate_timed": (_I, [_P, _U32, C.POINTER(_F)]),
    "blbm_collide": (_I, [_P]),
    "blbm_stream": (_I, [_P]),
    "blbm_set_summary": (_I, [_P, _I]),
    "blbm_rerender": (_I, [_P]),
    "blbm_compute_summary": (_I, [_P, _I]),
    "blbm_set_omega": (_I, [_P, _F]),
    "blbm_reset_to_equilibrium": (_I, [_P]),
    "blbm_custom_speed": (_I, [_P, _F]),
    "blbm_single_cell": (_I, [_P, _U32]),
    "blbm_draw_points": (_I, [_P, _P, _SZ]),
    "blbm_draw_points64": (_I, [_P, _P, _SZ]),
    "blbm_reset_barrier": (_I, [_P]),
    "blbm_write_barrier_rows": (_I, [_P, _U64, _U64, _P]),
    "blbm_timer_start": (_I, [_P]),
    "blbm_timer_stop": (_I, [_P, C.POINTER(_F)]),
    "blbm_rasterize_line": (_I, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _I, _P, _SZ,
                                 C.POINTER(_SZ)]),
    "blbm_preset_lines": (_I, [_I, C.c_int64, C.c_int64, _P, _SZ, C.POINTER(_SZ)]),
    "blbm_draw_line": (_I, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "blbm_erase_line": (_I, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "blbm_curl_barrier": (_I, [_P]),
    "blbm_chaos_barrier": (_I, [_P]),
    "blbm_welcome_barrier": (_I, [_P]),
    "blbm_get_compute_num": (_U64, [_P]),
    "blbm_get_frame_num": (_U64, [_P]),
    "blbm_read_population": (_I, [_P, _I, _I, _P]),
    "blbm_write_population": (_I, [_P, _I, _I, _P]),
    "blbm_read_moments": (_I, [_P, _P, _P, _P]),
    "blbm_read_output": (_I, [_P, _P]),
    "blbm_read_barrier": (_I, [_P, _P]),
    "blbm_read_cell_class": (_I, [_P, _P]),
    "blbm_color_map": (_I, [_P, _I]),
    "blbm_read_colors": (_I, [_P, _P]),
    "blbm_read_output_async": (_I, [_P, _P]),
    "blbm_synchronize": (_I, [_P]),
    "blbm_reduce_moments": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                 C.POINTER(_F)]),
    "blbm_get_geometry": (_I, [_P, C.POINTER(_U32), C.POINTER(_U64), C.POINTER(_U64), C.POINTER(_U64),
                               C.POINTER(_I)]),
    "blbm_link_local": (_I, [_P, _P]),
    "blbm_export_peer": (_I, [_P, _P]),
    "blbm_link_peer": (_I, [_P, _I, _P]),
    "blbm_exchange_halos": (_I, [_P]),
    "blbm_set_kernel": (_I, [_P, _I]),
    "blbm_get_kernel": (_I, [_P]),
    "blbm_set_tuning": (_I, [_P, _I, _I]),
    "blbm_set_lazy_barriers": (_I, [_P, _I]),
    "blbm_get_lazy_barriers_active": (_I, [_P]),
    "blbm_get_launch_count": (_U64, [_P]),
    "blbm_get_device_bytes": (_U64, [_P]),
}


def load_library():
    """dlopen libblbm.so and attach prototypes.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise BlbmError(-4, f"{path} not found: build it with `python -m lbm_b200.build` "
                                "(there is no CPU fallback)")
        lib = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        raise BlbmError(rc, load_library().blbm_last_error().decode(errors="replace"))


class LBM:
    """A D2Q9 BGK lattice: the whole lattice on one B200, one y-slab of it (`rows`), or — `devices=[...]` — the
    whole lattice cut into y-slabs over several B200s behind one handle (blbm_create_group)."""

    def __init__(self, omega, x, y, inflow_ux=0.1, device=0, rows=None, kernel=Kernel.Auto, lazy_barriers=None,
                 devices=None):
        self._L = load_library()
        self._h = _P()
        self.x, self.y = int(x), int(y)
        if devices is not None:
            if rows is not None:
                raise ValueError("rows= and devices= are mutually exclusive")
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            self.row_begin, self.row_end = 0, self.y
            self.device = int(devices[0]) if len(devices) else 0
            _check(self._L.blbm_create_group(self.x, self.y, float(omega), float(inflow_ux), devs, len(devices),
                                             C.byref(self._h)))
        else:
            if rows is None:
                rows = (0, self.y)
            self.row_begin, self.row_end = int(rows[0]), int(rows[1])
            self.device = int(device)
            _check(self._L.blbm_create_slab(self.x, self.y, self.row_begin, self.row_end, float(omega),
                                            float(inflow_ux), self.device, C.byref(self._h)))
        self.summary_stat = SummaryStat.Curl
        if kernel != Kernel.Auto:
            self.set_kernel(kernel)
        if lazy_barriers is not None:
            self.set_lazy_barriers(lazy_barriers)

    # -- life cycle
    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_owned", True):
                self._L.blbm_destroy(self._h)
            self._h = None

    def group_size(self):
        """number of slabs behind the handle (1 unless created with devices=[...])"""
        return int(self._L.blbm_group_size(self._h))

    def slab(self, index):
        """Non-owning view of slab `index` of a group handle (per-slab geometry, mask windows, tuning).  Calls that
        step or exchange halos belong on the group, which keeps its slabs in lock-step."""
        h = _P()
        _check(self._L.blbm_group_slab(self._h, int(index), C.byref(h)))
        v = object.__new__(LBM)
        v._L, v._h, v._owned = self._L, h, False
        w, hg, a, b, dev = _U32(), _U64(), _U64(), _U64(), _I()
        _check(self._L.blbm_get_geometry(h, C.byref(w), C.byref(hg), C.byref(a), C.byref(b), C.byref(dev)))
        v.x, v.y, v.row_begin, v.row_end, v.device = w.value, hg.value, a.value, b.value, dev.value
        v.summary_stat = self.summary_stat
        return v

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def local_rows(self):
        return self.row_end - self.row_begin

    # -- the reference's methods
    def iterate(self, compute_steps):
        _check(self._L.blbm_iterate(self._h, int(compute_steps)))

    def advance(self, compute_steps):
        _check(self._L.blbm_advance(self._h, int(compute_steps)))

    def iterate_timed(self, compute_steps):
        ms = _F()
        _check(self._L.blbm_iterate_timed(self._h, int(compute_steps), C.byref(ms)))
        return ms.value

    def collide(self):
        _check(self._L.blbm_collide(self._h))

    def stream(self):
        _check(self._L.blbm_stream(self._h))

    def rerender(self):
        _check(self._L.blbm_rerender(self._h))

    def set_summary(self, stat):
        _check(self._L.blbm_set_summary(self._h, int(stat)))
        self.summary_stat = SummaryStat(int(stat))

    def compute_summary(self, stat):
        _check(self._L.blbm_compute_summary(self._h, int(stat)))
        self.summary_stat = SummaryStat(int(stat))

    def update_omega_buffer(self, omega):
        _check(self._L.blbm_set_omega(self._h, float(omega)))

    def reset_to_equilibrium(self):
        _check(self._L.blbm_reset_to_equilibrium(self._h))

    def custom_speed(self, ux):
        _check(self._L.blbm_custom_speed(self._h, float(ux)))

    def single_cell(self, index):
        _check(self._L.blbm_single_cell(self._h, int(index)))

    def draw_points(self, pairs):
        """pairs: flat [loc, val, loc, val, ...] (merge_shapes.rs:12-22) or an (n, 2) array."""
        a = np.ascontiguousarray(pairs)
        if a.dtype == np.uint64 or (a.size and int(a.max()) > 0xFFFFFFFF):
            a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
            _check(self._L.blbm_draw_points64(self._h, a.ctypes.data, a.size // 2))
        else:
            a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1)
            _check(self._L.blbm_draw_points(self._h, a.ctypes.data, a.size // 2))

    def draw_shape(self, shape):
        """shape: anything with get_points() -> iterable of (x, y, bool), like `trait Shape`
        (barrier_shapes/mod.rs:11-19); flattened as get_points_vector does (merge_shapes.rs:12-22)."""
        a = points_vector(shape.get_points(), self.x)
        if not len(a):
            return  # the reference's callers guard with is_empty() (lib.rs:147)
        self.draw_points(a)

    def reset_barrier(self):
        _check(self._L.blbm_reset_barrier(self._h))

    def draw_line(self, p1, p2):
        """draw_shape(&Line::new(p1, p2, x, y)) — the thick Bresenham line of barrier_shapes/line.rs"""
        _check(self._L.blbm_draw_line(self._h, int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1])))

    def erase_line(self, p1, p2):
        _check(self._L.blbm_erase_line(self._h, int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1])))

    def curl_barrier(self):
        _check(self._L.blbm_curl_barrier(self._h))

    def chaos_barrier(self):
        _check(self._L.blbm_chaos_barrier(self._h))

    def welcome_barrier(self):
        _check(self._L.blbm_welcome_barrier(self._h))

    def write_barrier_rows(self, row_begin, mask_rows):
        """mask_rows: (nrows, W) uint8, 1 = barrier, for global rows [row_begin, row_begin+nrows)."""
        a = np.ascontiguousarray(mask_rows, dtype=np.uint8)
        assert a.ndim == 2 and a.shape[1] == self.x
        _check(self._L.blbm_write_barrier_rows(self._h, int(row_begin), a.shape[0], a.ctypes.data))

    def timer_start(self):
        _check(self._L.blbm_timer_start(self._h))

    def timer_stop(self):
        ms = _F()
        _check(self._L.blbm_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def get_compute_num(self):
        return int(self._L.blbm_get_compute_num(self._h))

    def get_frame_num(self):
        return int(self._L.blbm_get_frame_num(self._h))

    # -- read-back (new)
    def _shape(self):
        return (self.local_rows, self.x)

    def read_population(self, k, buffer=-1):
        out = np.empty(self._shape(), np.float32)
        _check(self._L.blbm_read_population(self._h, int(buffer), int(k), out.ctypes.data))
        return out

    def write_population(self, k, values, buffer=-1):
        a = np.ascontiguousarray(values, dtype=np.float32)
        assert a.shape == self._shape()
        _check(self._L.blbm_write_population(self._h, int(buffer), int(k), a.ctypes.data))

    def read_moments(self):
        mx, my, rho = (np.empty(self._shape(), np.float32) for _ in range(3))
        _check(self._L.blbm_read_moments(self._h, mx.ctypes.data, my.ctypes.data, rho.ctypes.data))
        return mx, my, rho

    def read_density(self):
        """the density plane alone (blbm_read_moments with the two momentum pointers NULL)"""
        rho = np.empty(self._shape(), np.float32)
        _check(self._L.blbm_read_moments(self._h, None, None, rho.ctypes.data))
        return rho

    def read_output(self):
        out = np.empty(self._shape(), np.float32)
        _check(self._L.blbm_read_output(self._h, out.ctypes.data))
        return out

    def color_map(self, cmap):
        """LBM::color_map (lbm.rs:1299): LUT of the output field into RGB on the device."""
        _check(self._L.blbm_color_map(self._h, int(cmap)))

    def read_colors(self):
        out = np.empty(self._shape() + (3,), np.float32)
        _check(self._L.blbm_read_colors(self._h, out.ctypes.data))
        return out

    def read_output_async(self, pinned_ptr):
        _check(self._L.blbm_read_output_async(self._h, int(pinned_ptr)))

    def read_barrier(self):
        out = np.empty(self._shape(), np.uint32)
        _check(self._L.blbm_read_barrier(self._h, out.ctypes.data))
        return out

    def read_cell_class(self):
        out = np.empty(self._shape(), np.uint16)
        _check(self._L.blbm_read_cell_class(self._h, out.ctypes.data))
        return out

    def reduce_moments(self):
        a, b, c, m = C.c_double(), C.c_double(), C.c_double(), _F()
        _check(self._L.blbm_reduce_moments(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(m)))
        return a.value, b.value, c.value, m.value

    def synchronize(self):
        _check(self._L.blbm_synchronize(self._h))

    # -- slabs
    def export_peer(self):
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        _check(self._L.blbm_export_peer(self._h, buf))
        return buf.raw

    def link_peer(self, side, blob):
        assert len(blob) == PEER_HANDLE_BYTES
        _check(self._L.blbm_link_peer(self._h, int(side), C.c_char_p(blob)))

    def exchange_halos(self):
        _check(self._L.blbm_exchange_halos(self._h))

    # -- tuning / measurement
    def set_kernel(self, kernel):
        _check(self._L.blbm_set_kernel(self._h, int(kernel)))

    def get_kernel(self):
        return Kernel(self._L.blbm_get_kernel(self._h))

    def set_tuning(self, knob, value):
        _check(self._L.blbm_set_tuning(self._h, int(knob), int(value)))

    def set_lazy_barriers(self, mode):
        """0 never, 1 always, 2 auto: keep barrier cells in the compact chain table (bit-identical)."""
        _check(self._L.blbm_set_lazy_barriers(self._h, int(mode)))

    def lazy_barriers_active(self):
        return bool(self._L.blbm_get_lazy_barriers_active(self._h))

    def launch_count(self):
        return int(self._L.blbm_get_launch_count(self._h))

    def device_bytes(self):
        return int(self._L.blbm_get_device_bytes(self._h))


def points_vector(points, xdim):
    """get_points_vector (merge_shapes.rs:12-22): (x, y, bool) -> [x + y*xdim, 1 | 0] pairs, the array
    draw_barrier_updates uploads for barrier_draw.wgsl (lbm.rs:1341-1343) and blbm_draw_points takes."""
    pts = list(points)
    a = np.empty((len(pts), 2), np.uint64)
    for q, (px, py, on) in enumerate(pts):
        a[q, 0] = int(px) + int(py) * int(xdim)
        a[q, 1] = 1 if on else 0
    return a


def rasterize_line(p1, p2, xdim, ydim, erase=False):
    """Cells of Line::new / Line::new_erased as an (n, 2) int64 array of (x, y); None for invalid end points."""
    L = load_library()
    n = _SZ()
    rc = L.blbm_rasterize_line(int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1]), int(xdim), int(ydim), int(erase), None,
                               0, C.byref(n))
    if rc != 0:
        return None
    out = np.empty((n.value, 2), np.int64)
    L.blbm_rasterize_line(int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1]), int(xdim), int(ydim), int(erase),
                          out.ctypes.data, n.value, C.byref(n))
    return out


def preset_lines(preset, xdim, ydim):
    """End points of the thick lines of a barrier preset (0 curl, 1 chaos, 2 welcome; lbm.rs:1367-1480) as an (n, 4)
    int64 array of (x1, y1, x2, y2), in the order the reference creates them."""
    L = load_library()
    n = _SZ()
    _check(L.blbm_preset_lines(int(preset), int(xdim), int(ydim), None, 0, C.byref(n)))
    out = np.empty((n.value, 4), np.int64)
    _check(L.blbm_preset_lines(int(preset), int(xdim), int(ydim), out.ctypes.data, n.value, C.byref(n)))
    return out


def slab_rows(y, nslabs):
    """Row ranges of the y-slab decomposition: contiguous, sizes differ by at most one row."""
    base, extra = divmod(int(y), int(nslabs))
    out, r = [], 0
    for s in range(nslabs):
        n = base + (1 if s < extra else 0)
        out.append((r, r + n))
        r += n
    return out


class SlabGroup(LBM):
    """A lattice split into y-slabs over several GPUs of THIS process: `LBM(devices=[...])`, i.e. one
    blbm_create_group handle — the slab orchestration lives behind the C ABI.  (One-process-per-GPU deployments
    link `LBM(rows=...)` slabs with export_peer/link_peer instead; see bench.py.)"""

    def __init__(self, omega, x, y, devices, inflow_ux=0.1, kernel=Kernel.Auto, lazy_barriers=None):
        super().__init__(omega, x, y, inflow_ux=inflow_ux, kernel=kernel, lazy_barriers=lazy_barriers,
                         devices=list(devices))
        self.ranges = slab_rows(y, len(devices))
