"""World-size-2 (and 3) CPU test of the N>1 path: the y-slab decomposition and its halo protocol.

Each gloo rank runs a numpy model of exactly what a GPU slab does: the "pull-T" step
on its own rows with one halo row above and two below, sending per step
    up:   rows n, ne, nw of its first row, w[x=0] of its first row, nw[x=0] of its second row
    down: rows s, se, sw of its last row
and nothing else; the barrier mask is replicated from the global paint list.  The concatenated slabs must be
bit-identical to the undivided oracle.  This validates the protocol (which rows, which populations, the
flat-index wrap at column W-1 needing column 0 of rows y+1 and y+2) without a GPU; the CUDA implementation
of the same protocol is covered by the -m gpu slab tests.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

F = np.float32
# moving populations in device order: nw n ne w e sw s se, travel vectors (dx, dy), y down
DIRS = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]
D_NW, D_N, D_NE, D_W, D_E, D_SW, D_S, D_SE = range(8)
POP_OF_DIR = [0, 1, 2, 3, 5, 6, 7, 8]


def collide(f, rest, omega):
    """Vectorised fp32 BGK collision, op order of collision/*.wgsl (same as oracle/lbm_numpy.py)."""
    nw, n, ne, w, e, sw, s, se = f
    om = F(omega)
    one = F(1.0)
    with np.errstate(all="ignore"):
        mx = ne + se - nw - sw
        my = ne + nw - se - sw
        rho = ne + se + nw + sw
        mx = mx + (e - w)
        my = my + (n - s)
        rho = rho + (e + n + s + w)
        rho = rho + rest
        ux, uy = mx / rho, my / rho
        k36, k9 = F(1.0 / 36.0) * rho, F(1.0 / 9.0) * rho
        ux3, uy3 = F(3.0) * ux, F(3.0) * uy
        ux2, uy2 = ux * ux, uy * uy
        uxuy2 = F(2.0) * ux * uy
        u2 = ux2 + uy2
        u215 = F(1.5) * u2
        out = [None] * 8
        out[D_NE] = ne + om * (k36 * (one + ux3 + uy3 + F(4.5) * (u2 + uxuy2) - u215) - ne)
        out[D_SE] = se + om * (k36 * (one + ux3 - uy3 + F(4.5) * (u2 - uxuy2) - u215) - se)
        out[D_NW] = nw + om * (k36 * (one - ux3 + uy3 + F(4.5) * (u2 - uxuy2) - u215) - nw)
        out[D_SW] = sw + om * (k36 * (one - ux3 - uy3 + F(4.5) * (u2 + uxuy2) - u215) - sw)
        rest2 = rest + om * (F(4.0 / 9.0) * rho * (one - u215) - rest)
        out[D_E] = e + om * (k9 * (one + ux3 + F(4.5) * ux2 - u215) - e)
        out[D_W] = w + om * (k9 * (one - ux3 + F(4.5) * ux2 - u215) - w)
        out[D_N] = n + om * (k9 * (one + uy3 + F(4.5) * uy2 - u215) - n)
        out[D_S] = s + om * (k9 * (one - uy3 + F(4.5) * uy2 - u215) - s)
    return out, rest2, mx, my, rho


class SlabModel:
    """Rows [r0, r1) of a W x H lattice; planes have rows+3 rows (1 halo above, 2 below), the mask rows+4."""

    def __init__(self, w, h, r0, r1, omega, u0, equil):
        self.w, self.h, self.r0, self.r1, self.rows = w, h, r0, r1, r1 - r0
        self.omega = omega
        self.f = [[np.zeros((self.rows + 3, w), F) for _ in range(8)] for _ in range(2)]
        self.rest = np.zeros((self.rows + 3, w), F)
        lo = 0 if r0 > 0 else 1
        hi = self.rows + 1 + (min(2, h - r1) if r1 < h else 0)
        for b in range(2):
            for d in range(8):
                self.f[b][d][lo:hi] = equil[POP_OF_DIR[d]]
        self.rest[lo:hi] = equil[4]
        self.mask = np.zeros((self.rows + 4, w), np.uint8)  # mask row m <-> global row r0 - 2 + m
        for m in range(self.rows + 4):
            gy = r0 - 2 + m
            if gy == 0 or gy == h - 1:
                self.mask[m] = 1
        self.step = 0
        self.regime_t = False

    def mask_at(self, gx, gy):
        """flat-index neighbour lookup with out-of-range -> 0 (vectorised over arrays gx, gy)."""
        gx, gy = gx.copy(), gy.copy()
        wrap_e = gx == self.w
        gx[wrap_e] = 0
        gy[wrap_e] += 1
        wrap_w = gx < 0
        gx[wrap_w] = self.w - 1
        gy[wrap_w] -= 1
        ok = (gy >= 0) & (gy < self.h)
        m = np.clip(gy - self.r0 + 2, 0, self.rows + 3)
        return np.where(ok, self.mask[m, gx], 0)

    def draw(self, pairs):
        for loc, val in np.asarray(pairs).reshape(-1, 2):
            gy, gx = divmod(int(loc), self.w)
            if gy >= self.h:
                continue
            m = gy - self.r0 + 2
            if 0 <= m < self.rows + 4:
                self.mask[m, gx] = 1 if val == 1 else 0

    def _gather(self, x_buf):
        """S_k on own rows from T_{k-1} in buffer x_buf; returns (list of 8 arrays, active mask)."""
        w, rows = self.w, self.rows
        ys, xs = np.mgrid[0:rows, 0:w]
        gy = ys + self.r0
        bar = self.mask[2:2 + rows] == 1
        active = ~bar & (xs != 0) & (gy < self.h - 1)
        out = []
        for d, (dx, dy) in enumerate(DIRS):
            sx, sy = xs - dx, ys - dy          # geometric source (local row index, may be -1 .. rows+1)
            wrap = sx == w                     # column W-1 pulling a west-moving population: flat index
            sx = np.where(wrap, 0, sx)
            sy = np.where(wrap, sy + 1, sy)
            sxc = np.clip(sx, 0, w - 1)        # sx = -1 only at x = 0, which is never active
            pulled = self.f[x_buf][d][sy + 1, sxc]
            up_is_bar = self.mask_at(xs - dx, gy - dy) == 1
            own_opp = self.f[x_buf][7 - d][1:1 + rows]
            out.append(np.where(up_is_bar, own_opp, pulled))
        return out, active

    def step_fused_or_collide(self):
        y = self.step % 2
        own = slice(1, 1 + self.rows)
        if not self.regime_t:
            f = [self.f[y][d][own] for d in range(8)]
            self.regime_t = True
        else:
            g, active = self._gather(1 - y)
            f = [np.where(active, g[d], self.f[y][d][own]) for d in range(8)]
        newf, rest2, mx, my, rho = collide(f, self.rest[own], self.omega)
        for d in range(8):
            self.f[y][d][own] = newf[d]
        self.rest[own] = rest2
        self.step += 1
        return y

    def materialise(self):
        if not self.regime_t:
            return
        y = self.step % 2
        g, active = self._gather(1 - y)
        own = slice(1, 1 + self.rows)
        for d in range(8):
            self.f[y][d][own] = np.where(active, g[d], self.f[y][d][own])
        self.regime_t = False


def exchange(slab, y_buf, rank, world):
    """The per-step halo protocol over torch.distributed (gloo): exactly the data the CUDA kernel stores
    into its neighbours' halo rows."""
    w, rows = slab.w, slab.rows
    reqs, recv_up, recv_dn = [], None, None
    if rank > 0:  # send up
        first = np.concatenate([slab.f[y_buf][d][1] for d in (D_N, D_NE, D_NW)] +
                               [slab.f[y_buf][D_W][1, :1], slab.f[y_buf][D_NW][2, :1]]).copy()
        reqs.append(dist.isend(torch.from_numpy(first), rank - 1))
        recv_up = torch.empty(3 * w, dtype=torch.float32)
        reqs.append(dist.irecv(recv_up, rank - 1))
    if rank < world - 1:  # send down
        last = np.concatenate([slab.f[y_buf][d][rows] for d in (D_S, D_SE, D_SW)]).copy()
        reqs.append(dist.isend(torch.from_numpy(last), rank + 1))
        recv_dn = torch.empty(3 * w + 2, dtype=torch.float32)
        reqs.append(dist.irecv(recv_dn, rank + 1))
    for r in reqs:
        r.wait()
    if recv_up is not None:
        a = recv_up.numpy()
        for q, d in enumerate((D_S, D_SE, D_SW)):
            slab.f[y_buf][d][0] = a[q * w:(q + 1) * w]
    if recv_dn is not None:
        a = recv_dn.numpy()
        for q, d in enumerate((D_N, D_NE, D_NW)):
            slab.f[y_buf][d][rows + 1] = a[q * w:(q + 1) * w]
        slab.f[y_buf][D_W][rows + 1, 0] = a[3 * w]
        slab.f[y_buf][D_NW][rows + 2, 0] = a[3 * w + 1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, steps_a, steps_b, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lbm_b200.lbm import slab_rows
        from oracle import lbm_numpy
        r0, r1 = slab_rows(h, world)[rank]
        omega = 1.0 / (3 * 0.02 + 0.5)
        equil = lbm_numpy.set_equil(0.1, 0.0, 1.0)
        slab = SlabModel(w, h, r0, r1, omega, 0.1, equil)
        # paints straddling every slab boundary, columns 0 / 1 / W-2 / W-1 included, plus an obstacle
        bounds = [r[0] for r in slab_rows(h, world)[1:]]
        loc = [y * w + x for b in bounds for y in (b - 2, b - 1, b, b + 1) for x in (0, 1, w // 2, w - 2, w - 1)]
        loc += [(h // 2 + dy) * w + (w // 3 + dx) for dy in range(-2, 3) for dx in range(-2, 3)]
        pairs = np.array([[l, 1] for l in loc], np.uint32)
        slab.draw(pairs)
        for _ in range(steps_a):
            y = slab.step_fused_or_collide()
            exchange(slab, y, rank, world)
        # erase part of it mid-run (mask change takes effect like the reference: after the pending stream)
        slab.materialise()
        erase = np.array([[l, 0] for l in loc[::3]], np.uint32)
        slab.draw(erase)
        for _ in range(steps_b):
            y = slab.step_fused_or_collide()
            exchange(slab, y, rank, world)
        slab.materialise()
        live = slab.step % 2
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), r0=r0, r1=r1,
                 **{f"f{b}_{d}": slab.f[b][d][1:1 + slab.rows] for b in range(2) for d in range(8)},
                 rest=slab.rest[1:1 + slab.rows], live=live)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_halo_protocol_matches_undivided_oracle(world, tmp_path):
    from oracle.lbm_oracle import Oracle
    from lbm_b200.lbm import slab_rows
    w, h, steps_a, steps_b = 48, 26, 37, 24
    port = _free_port()
    mp.spawn(_worker, args=(world, port, w, h, steps_a, steps_b, str(tmp_path)), nprocs=world, join=True)
    # the same scenario on the undivided oracle
    o = Oracle(1.0 / (3 * 0.02 + 0.5), w, h)
    bounds = [r[0] for r in slab_rows(h, world)[1:]]
    loc = [y * w + x for b in bounds for y in (b - 2, b - 1, b, b + 1) for x in (0, 1, w // 2, w - 2, w - 1)]
    loc += [(h // 2 + dy) * w + (w // 3 + dx) for dy in range(-2, 3) for dx in range(-2, 3)]
    o.draw_points(np.array([[l, 1] for l in loc], np.uint32))
    o.iterate(steps_a)
    o.draw_points(np.array([[l, 0] for l in loc[::3]], np.uint32))
    o.iterate(steps_b)
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    for b in range(2):
        for d in range(8):
            got = np.concatenate([p[f"f{b}_{d}"] for p in parts], axis=0)
            want = o.population(b, POP_OF_DIR[d])
            bad = (got.view(np.uint32) != want.view(np.uint32)) & ~(np.isnan(got) & np.isnan(want))
            assert not bad.any(), f"buffer {b} dir {d}: {bad.sum()} cells differ, first {np.argwhere(bad)[0]}"
    got = np.concatenate([p["rest"] for p in parts], axis=0)
    assert (got.view(np.uint32) == o.population(0, 4).view(np.uint32)).all()
