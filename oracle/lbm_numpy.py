"""Second, independent CPU restatement of the lbm-wgpu lattice update in vectorised numpy.

TEST INFRASTRUCTURE ONLY (see oracle/lbm_oracle.c).  Purpose: a different code path (whole-array
slicing on zero-padded flat arrays instead of per-cell loops) that must agree BIT-FOR-BIT with the C
oracle; a transcription slip in either shows up as a mismatch.  Parity status as for the C oracle: pinned
to the reference's executed WGSL text and to its wasm binary's host functions, not to a WebGPU-backend run.

numpy float32 arithmetic rounds every binary op individually (no FMA), `/` is IEEE division — the
oracle semantics of SURVEY.md section 8.  Reference files: the same list as lbm_oracle.c
(lbm-wgpu/src/lbm.rs:595-643, :1065-1134, :1174-1252; rewritten_shaders/{pre_collision,collision,
stream,summary_stats,update_barrier}/*.wgsl).
"""
import numpy as np

F = np.float32
NW, N, NE, W, REST, E, SW, S, SE = range(9)


def set_equil(ux, uy, rho):
    """lbm.rs:611-643, fp32 scalar arithmetic in the order written."""
    ux, uy, rho = F(ux), F(uy), F(rho)
    ux2 = ux * ux
    uy2 = uy * uy
    ud = ux2 + uy2
    pos = ud + F(2.0) * (ux * uy)
    neg = ud - F(2.0) * (ux * uy)
    ux = ux * F(3.0)
    uy = uy * F(3.0)
    ux2 = ux2 * F(4.5)
    uy2 = uy2 * F(4.5)
    ud = ud * F(1.5)
    neg = neg * F(4.5)
    pos = pos * F(4.5)
    r9 = rho / F(9.0)
    r36 = rho / F(36.0)
    one = F(1.0)
    return np.array([
        r36 * (one - ux + uy + neg - ud),
        r9 * (one + uy + uy2 - ud),
        r36 * (one + ux + uy + pos - ud),
        r9 * (one - ux + ux2 - ud),
        F(4.0) * r9 * (one - ud),
        r9 * (one + ux + ux2 - ud),
        r36 * (one - ux - uy + pos - ud),
        r9 * (one - uy - uy2 - ud),
        r36 * (one + ux - uy + neg - ud),
    ], dtype=F)


class NumpyLBM:
    """Flat arrays of length W*H, index i = x + y*W (lbm.rs:607-609)."""

    def __init__(self, omega, x, y, inflow_ux=0.1):
        self.w, self.h = int(x), int(y)
        self.n = self.w * self.h
        v = set_equil(inflow_ux, 0.0, 1.0)
        self.f = [[np.full(self.n, v[k], F) for k in range(9)] for _ in range(2)]
        self.bar = self._init_barrier()
        self.mx = np.zeros(self.n, F)
        self.my = np.zeros(self.n, F)
        self.rho = np.zeros(self.n, F)
        self.out = np.zeros(self.n, F)
        self.omega = F(omega)
        self.step_no = 0
        self.stat = 0

    def _init_barrier(self):
        b = np.zeros(self.n, np.uint32)
        b[: self.w] = 1
        b[(self.h - 1) * self.w:] = 1
        return b

    # shifted read with robust-buffer-access semantics: result[i] = a[i + off], 0 outside [0, n)
    def _sh(self, a, off):
        pad = self.w + 2
        p = np.zeros(self.n + 2 * pad, a.dtype)
        p[pad:pad + self.n] = a
        return p[pad + off: pad + off + self.n]

    def _pre_collision(self, c):
        f = self.f[c]
        self.mx = f[NE] + f[SE] - f[NW] - f[SW]
        self.my = f[NE] + f[NW] - f[SE] - f[SW]
        self.rho = f[NE] + f[SE] + f[NW] + f[SW]
        self.mx = self.mx + (f[E] - f[W])
        self.my = self.my + (f[N] - f[S])
        self.rho = self.rho + (f[E] + f[N] + f[S] + f[W])

    def collide(self):
        c = self.step_no % 2
        f = self.f[c]
        om = self.omega
        self._pre_collision(c)
        origin = self.f[0][REST]
        with np.errstate(all="ignore"):
            # corner_collision.wgsl
            self.rho = self.rho + origin
            rho = self.rho
            ux = self.mx / rho
            uy = self.my / rho
            k36 = F(1.0 / 36.0) * rho
            ux3 = F(3.0) * ux
            uy3 = F(3.0) * uy
            ux2 = ux * ux
            uy2 = uy * uy
            uxuy2 = F(2.0) * ux * uy
            u2 = ux2 + uy2
            u215 = F(1.5) * u2
            one = F(1.0)
            f[NE] = f[NE] + om * (k36 * (one + ux3 + uy3 + F(4.5) * (u2 + uxuy2) - u215) - f[NE])
            f[SE] = f[SE] + om * (k36 * (one + ux3 - uy3 + F(4.5) * (u2 - uxuy2) - u215) - f[SE])
            f[NW] = f[NW] + om * (k36 * (one - ux3 + uy3 + F(4.5) * (u2 - uxuy2) - u215) - f[NW])
            f[SW] = f[SW] + om * (k36 * (one - ux3 - uy3 + F(4.5) * (u2 + uxuy2) - u215) - f[SW])
            # cardinal_collision.wgsl
            k9 = F(1.0 / 9.0) * rho
            self.f[0][REST] = origin + om * (F(4.0 / 9.0) * rho * (one - u215) - origin)
            f[E] = f[E] + om * (k9 * (one + ux3 + F(4.5) * ux2 - u215) - f[E])
            f[W] = f[W] + om * (k9 * (one - ux3 + F(4.5) * ux2 - u215) - f[W])
            f[N] = f[N] + om * (k9 * (one + uy3 + F(4.5) * uy2 - u215) - f[N])
            f[S] = f[S] + om * (k9 * (one - uy3 + F(4.5) * uy2 - u215) - f[S])

    def stream(self):
        c = self.step_no % 2
        d = 1 - c
        i = np.arange(self.n, dtype=np.int64)
        active = (self.bar != 1) & (i % self.w != 0) & (i // self.w < self.h - 1)
        row = self.w
        for a, b, off in ((E, W, 1), (S, N, row), (SE, NW, 1 + row), (NE, SW, 1 - row)):
            src_a, src_b = self.f[c][a], self.f[c][b]
            new_b = np.where(self._sh(self.bar, off) == 1, src_a, self._sh(src_b, off))
            new_a = np.where(self._sh(self.bar, -off) == 1, src_b, self._sh(src_a, -off))
            self.f[d][b] = np.where(active, new_b, self.f[d][b])
            self.f[d][a] = np.where(active, new_a, self.f[d][a])

    def step(self):
        self.collide()
        self.stream()
        self.step_no += 1

    def summary(self, stat):
        self.stat = stat
        mx, my, rho = self.mx, self.my, self.rho
        with np.errstate(all="ignore"):
            if stat == 0:
                i = np.arange(self.n, dtype=np.int64)
                m = (i % self.w != 0) & (i // self.w < self.h - 1)
                v = F(10.0) * (self._sh(my, 1) - self._sh(my, -1) - self._sh(mx, -self.w)
                               + self._sh(mx, self.w)) / rho
                self.out = np.where(m, v, self.out).astype(F)
            elif stat == 1:
                self.out = mx.copy()
            elif stat == 2:
                self.out = my.copy()
            elif stat == 3:
                self.out = F(4.0) * np.minimum(np.maximum(F(0.15) * rho, F(0)), F(1)) - F(0.5)
            elif stat == 4:
                self.out = np.minimum(np.maximum(F(5.0) * np.sqrt(mx * mx + my * my), F(0)), F(1)) - F(0.5)

    def iterate(self, n):
        for _ in range(n):
            self.step()
        self.summary(self.stat)

    def draw_points(self, pairs):
        p = np.asarray(pairs, dtype=np.uint32).reshape(-1, 2)
        for loc, val in p:  # sequential: last writer wins, like a serialised scatter
            if loc < self.n:
                self.bar[loc] = val

    def reset_barrier(self):
        self.bar = self._init_barrier()

    def update_omega_buffer(self, omega):
        self.omega = F(omega)

    def custom_speed(self, ux):
        v = set_equil(ux, 0.0, 1.0)
        self.f = [[np.full(self.n, v[k], F) for k in range(9)] for _ in range(2)]
        self.step_no = 0
        self._pre_collision(0)

    def reset_to_equilibrium(self):
        self.custom_speed(0.1)

    def single_cell(self, index):
        v = set_equil(0.0, 0.0, 1.0)
        self.f = [[np.full(self.n, v[k], F) for k in range(9)] for _ in range(2)]
        x, y = self.w, self.h
        cells = {0: (x - 2, y - 2), 1: (3 * x // 4, y - 2), 2: (x // 3, y - 2), 3: (x - 2, y // 2),
                 4: (3 * x // 4, y // 2), 5: (x // 2, y // 2), 6: (x - 2, 1), 7: (3 * x // 4, 1),
                 8: (x // 2, 1)}
        if index in cells:
            cx, cy = cells[index]
            i = cx + cy * x
            if 0 <= i < self.n:
                self.f[0][index][i] = F(4.0)
                self.f[1][index][i] = F(4.0)
        self.step_no = 0

    def population(self, buffer, k):
        if buffer < 0:
            buffer = self.step_no % 2
        if k == REST:
            buffer = 0
        return self.f[buffer][k].reshape(self.h, self.w)
