"""Generates tests/golden/wasm_golden.npz by EXECUTING HOST-SIDE FUNCTIONS OF THE REFERENCE'S OWN SHIPPED BINARY
(/root/reference/lbm-wgpu/pkg/lbm_wgpu_bg.wasm, the wasm-pack build of lbm-wgpu) in oracle/wasm_mini.py:

  * `set_equil` (lbm.rs:611-643; rustc specialised it to uy = 0, rho = 1, the only values lbm.rs ever passes):
    the nine initial populations for a list of inflow speeds — row a-2 of SURVEY.md section 8;
  * `LBM::single_cell` (lbm.rs:1502-1515, with `set_single_cell` :1482-1500 inlined) on a fabricated `LBM` value,
    its 18 `queue.write_buffer` calls intercepted: which population is set to 4.0 at which cell for every preset
    index, and what every other cell holds — row a-8;
  * `LBM::draw_shape` (lbm.rs:1337-1343, with `get_points_vector`, merge_shapes.rs:12-22, inlined) on Curves holding
    drawn and erased lines: the exact u32 pairs and the count word it uploads for barrier_draw.wgsl — the wire format
    of row a-7, i.e. what blbm_draw_points must accept;
  * `LBM::iterate` (lbm.rs:1065-1074 with compute_step, collide, stream, calculate_summary, color_map inlined) against
    a mock of wgpu's `dyn DynContext`: the complete command stream — encoders, the eight labelled compute passes per
    step with their pipeline and bind-group fields, which bind groups alternate with compute_step % 2, the dispatch
    size field, the summary and colour-map passes — i.e. the host-side dispatch order of rows a-3 .. a-6;
  * `Line::new` (barrier_shapes/line.rs:22-54) called directly, and `Line::new_erased` (line.rs:56-87, the 30-wide
    eraser) and `Line::new` again through `Curve::erase_segment` / `Curve::add_segment` (curve.rs:28-48): the
    point sets of thick lines, with the un-vendored `line_drawing 1.0.0` Bresenham as compiled into the binary —
    row N2 (the rasteriser in front of draw_points).

  * the barrier presets `LBM::curl_barrier`, `chaos_barrier`, `welcome_barrier` (lbm.rs:1367-1480).  rustc inlined them
    into the event-loop closure (lib.rs:108-135, the `match *BARRIER_PRESET` arms), so they cannot be called; instead
    the interpreter EXECUTES THE ARM ITSELF: `Instance.run_fragment` starts at the arm's first instruction with the
    two locals it reads (`&mut lbm`, `&driver`) pointing at a fabricated LBM (xdim, ydim) and runs until control
    leaves the arm, with `Line::new` logged (its six integer arguments: the end-point arithmetic) and
    `LBM::draw_shape` served by the host (the `dyn Shape` it is handed is read out through its own vtable's
    `get_points`): every thick line of every preset and the exact point sets drawn, in order -> wasm_presets.npz.

The binary carries no name section, so the functions are addressed by index and each index is checked against a
fingerprint (signature, floating-point constants, the "Endpoints (" format string it refers to, its callees) before
it is trusted; the SHA-256 of the binary is stored with the results.  None of these functions touches an import on
its normal path (std's HashSet keys are constants on wasm32-unknown-unknown), so no browser is needed.

    python tests/golden/make_wasm_golden.py        # build container only (needs /root/reference); ~1 minute
"""
import hashlib
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.wasm_mini import Instance, Module, Trap  # noqa: E402

WASM = "/root/reference/lbm-wgpu/pkg/lbm_wgpu_bg.wasm"
F_SET_EQUIL, F_LINE_NEW, F_ERASE_SEGMENT, F_ADD_SEGMENT = 268, 249, 307, 308
F_SINGLE_CELL, F_WRITE_BUFFER, F_DRAW_SHAPE = 247, 591, 267
# struct LBM as laid out in this build (read off the code of LBM::single_cell): xdim, ydim, data_buffers {ptr, len};
# a wgpu::Buffer is 88 bytes, the wgpu::Queue sits 128 bytes into Driver
LBM_X, LBM_Y, LBM_BUFFERS_PTR, LBM_BUFFERS_LEN, SIZEOF_BUFFER, DRIVER_QUEUE = 1100, 1104, 1160, 1164, 88, 128
SINGLE_CELL_SIZES = ((16, 12), (64, 32), (30, 17), (9, 7))
# LBM::iterate against a mock of wgpu's `dyn DynContext`: vtable slots of this build, identified on draw_shape (whose
# call sequence is known from lbm.rs:1341-1356) — and more LBM fields, found by perturbing them
F_ITERATE = 246
# the barrier presets are arms of a `match` inside the event-loop closure (the one function that calls Line::new
# dozens of times); PRESET_SIZES are lattice sizes on which every end point is valid (Line::new(..).unwrap())
PRESET_SIZES = ((512, 256), (300, 170), (1000, 500), (130, 64), (1920, 1080))
SLOT = {"queue_write_buffer": 85, "queue_submit": 91, "device_create_command_encoder": 37, "begin_compute_pass": 72,
        "end_compute_pass": 73, "encoder_finish": 76, "set_pipeline": 96, "set_bind_group": 97,
        "dispatch_workgroups": 105}
LBM_COMPUTE_STEP, LBM_WORK_GROUP_SIZE, LBM_STAT_CMAP, DRIVER_DEVICE = 1088, 1096, 1168, 104
I32, F32 = 0x7F, 0x7D

INFLOWS = (0.1, 0.05, 0.0, 0.07, 0.03, 0.3, -0.1, 0.2)


def line_cases():
    """(p1, p2, xdim, ydim): every octant, axis-aligned, degenerate, border-to-border, and random pairs"""
    c = []
    for xd, yd in ((64, 32), (512, 256)):
        mx, my = xd - 1, yd - 1
        pairs = [((3, 4), (20, 11)), ((20, 11), (3, 4)), ((5, 5), (5, 5)), ((0, 0), (mx, my)), ((mx, 0), (0, my)),
                 ((10, 3), (10, my - 2)), ((2, 9), (mx - 3, 9)), ((7, my - 1), (8, 2)), ((0, 0), (0, 0)),
                 ((mx, my), (mx, my)), ((1, 1), (2, 2)), ((1, 2), (2, 1)), ((30, 3), (33, 28)), ((33, 28), (30, 3)),
                 ((4, 20), (40, 17)), ((40, 17), (4, 20)), ((0, my), (mx, my)), ((mx, 0), (mx, my))]
        rng = np.random.default_rng(1000 + xd)
        for _ in range(30):
            pairs.append(((int(rng.integers(0, xd)), int(rng.integers(0, yd))),
                          (int(rng.integers(0, xd)), int(rng.integers(0, yd)))))
        c += [(a, b, xd, yd) for a, b in pairs]
    # invalid end points: Line::new returns Err
    c += [((3, 4), (64, 11), 64, 32), ((-1, 4), (10, 11), 64, 32), ((3, 32), (10, 11), 64, 32)]
    return c


def curve_chains():
    """(ops, xdim, ydim) with ops = [(erase, x, y), ...]: strokes as the click handler builds them (lib.rs:130-160)"""
    rng = np.random.default_rng(77)
    chains = [([(0, 5, 5)], 64, 32), ([(1, 9, 9)], 64, 32),
              ([(0, 3, 4), (0, 20, 11), (0, 25, 30), (0, 60, 2)], 64, 32),
              ([(0, 10, 10), (0, 40, 10), (1, 40, 25), (0, 12, 25)], 64, 32),
              ([(1, 30, 16), (1, 34, 16)], 64, 32)]
    for _ in range(4):
        n = int(rng.integers(3, 7))
        chains.append(([(int(rng.integers(0, 4) == 0), int(rng.integers(0, 200)), int(rng.integers(0, 120)))
                        for _ in range(n)], 200, 120))
    return chains


def check_fingerprints(m):
    def consts(f, op):
        return {i[1] for i in m.decode(f)[1] if i[0] == op}

    def calls(f):
        return {i[1] for i in m.decode(f)[1] if i[0] == 0x10}

    endpoints = None
    for off, blob in m.segments:
        k = blob.find(b"Endpoints (")
        if k >= 0:
            endpoints = off + k
    assert endpoints is not None
    # the format-pieces table right behind the string starts with (ptr to "Endpoints (", 11)
    pieces = None
    for off, blob in m.segments:
        k = blob.find(struct.pack("<II", endpoints, 11))
        if k >= 0:
            pieces = off + k
    assert pieces is not None
    assert m.type_of(F_SET_EQUIL) == ([I32, F32, I32, I32], [])
    f = consts(F_SET_EQUIL, 0x43)
    assert {4.5, 1.5, 3.0, 1.0}.issubset(f)
    assert float(np.float32(1.0) / np.float32(36.0)) in f and float(np.float32(1.0) / np.float32(9.0)) in f
    assert m.type_of(F_LINE_NEW) == ([I32] * 7, []) and pieces in consts(F_LINE_NEW, 0x41)
    assert m.type_of(F_ERASE_SEGMENT) == ([I32] * 5, []) and pieces in consts(F_ERASE_SEGMENT, 0x41)
    assert m.type_of(F_ADD_SEGMENT) == ([I32] * 5, []) and F_LINE_NEW in calls(F_ADD_SEGMENT)
    assert F_LINE_NEW not in calls(F_ERASE_SEGMENT)  # new_erased is inlined there
    # single_cell(&mut self, &driver, index): one set_equil, then 18 write_buffer calls; 4.0 is stored as its bit pattern
    code = m.decode(F_SINGLE_CELL)[1]
    assert m.type_of(F_SINGLE_CELL) == ([I32] * 3, []) and m.type_of(F_WRITE_BUFFER) == ([I32] * 4, [])
    assert [i[1] for i in code if i[0] == 0x10].count(F_SET_EQUIL) == 1
    assert [i[1] for i in code if i[0] == 0x10].count(F_WRITE_BUFFER) >= 18
    assert 0x40800000 in consts(F_SINGLE_CELL, 0x41)
    # draw_shape(&mut self, &driver, &dyn Shape): one trait-object call (get_points), two uploads, reads self.x
    code = m.decode(F_DRAW_SHAPE)[1]
    assert m.type_of(F_DRAW_SHAPE) == ([I32] * 4, [])
    assert [i[1] for i in code if i[0] == 0x10].count(F_WRITE_BUFFER) == 2
    assert any(i[0] == 0x28 and i[1] == LBM_X for i in code) and code[[i[0] for i in code].index(0x11)][0] == 0x11
    # hashbrown's static empty control group, referenced where the functions create an empty HashSet
    empty = [c for c in consts(F_LINE_NEW, 0x41) if 1 << 20 <= c < 1 << 21 and
             any(off <= c < off + len(b) and b[c - off:c - off + 4] == b"\xff" * 4 for off, b in m.segments)]
    assert len(empty) == 1
    return empty[0]


class Reference:
    """the shipped binary, its located functions and the calling conventions found by inspection"""

    def __init__(self, path=WASM):
        self.blob = open(path, "rb").read()
        self.sha256 = hashlib.sha256(self.blob).hexdigest()
        self.m = Module(self.blob)
        self.empty_group = check_fingerprints(self.m)
        self.stubs = {n: (lambda inst, *a: 37) if self.m.types[t][1] else (lambda inst, *a: None)
                      for _, n, t in self.m.imports}

    def _instance(self, frame=64):
        inst = Instance(self.m, imports=self.stubs)
        base = inst.globals[0] - frame  # a frame on the shadow stack for the out-parameter
        inst.globals[0] = base
        return inst, base

    def set_equil(self, ux, x=5, y=3):
        """fn set_equil(ux, uy = 0, rho = 1, x, y) -> Vec<Vec<f32>>: nine vectors of x*y equal values"""
        inst, ret = self._instance()
        inst.call(F_SET_EQUIL, ret, float(np.float32(ux)), x, y)
        cap, ptr, n = inst.u32(ret), inst.u32(ret + 4), inst.u32(ret + 8)
        assert n == 9 and not inst.called
        out = np.zeros(9, np.float32)
        for k in range(9):
            p, ln = inst.u32(ptr + 12 * k + 4), inst.u32(ptr + 12 * k + 8)
            v = np.frombuffer(inst.read(p, 4 * ln), dtype=np.float32)
            assert ln == x * y and (v.view(np.uint32) == v.view(np.uint32)[0]).all()
            out[k] = v[0]
        return out

    @staticmethod
    def _read_set(inst, base):
        """HashSet<(isize, isize, bool)> = RandomState {k0, k1: u64} + RawTable {bucket_mask, growth_left, items, ctrl};
        12-byte elements lie below ctrl, element i at ctrl - 12 (i + 1); a control byte with the top bit clear = full"""
        mask, _, items, ctrl = (inst.u32(base + 16 + 4 * k) for k in range(4))
        if ctrl == 0:
            return None  # Err(String): the niche of the control pointer
        pts = []
        for i in range(mask + 1):
            if inst.mem[ctrl + i] & 0x80 == 0:
                a = ctrl - 12 * (i + 1)
                pts.append((inst.i32(a), inst.i32(a + 4), inst.mem[a + 8]))
        assert len(pts) == items == len(set(pts))
        return sorted(pts)

    def single_cell(self, index, x, y):
        """LBM::single_cell(&mut self, &driver, index) on an LBM whose only meaningful fields are xdim, ydim and
        data_buffers (2 x 9 dummy wgpu::Buffer values); returns data[b][k] = the array uploaded to data_buffers[b][k]"""
        got = {}
        inst = None

        def write_buffer(inst_, queue, buffer, ptr, nbytes):
            b, k = divmod((buffer - bufs) // SIZEOF_BUFFER, 9)
            assert queue == drv + DRIVER_QUEUE and (buffer - bufs) % SIZEOF_BUFFER == 0 and (b, k) not in got
            got[(b, k)] = np.frombuffer(inst_.read(ptr, nbytes), dtype=np.float32).copy()

        inst = Instance(self.m, imports=self.stubs, hooks={F_WRITE_BUFFER: write_buffer})
        malloc = self.m.exports["__wbindgen_malloc"][1]
        me = inst.call(malloc, 2048, 8)
        drv = inst.call(malloc, 1024, 8)
        bufs = inst.call(malloc, 18 * SIZEOF_BUFFER, 8)
        outer = inst.call(malloc, 24, 4)
        for a, n in ((me, 2048), (drv, 1024), (bufs, 18 * SIZEOF_BUFFER)):
            inst.mem[a:a + n] = bytes(n)
        struct.pack_into("<IIIIII", inst.mem, outer, 9, bufs, 9, 9, bufs + 9 * SIZEOF_BUFFER, 9)  # two Vec {cap, ptr, len}
        struct.pack_into("<II", inst.mem, me + LBM_X, x, y)
        struct.pack_into("<II", inst.mem, me + LBM_BUFFERS_PTR, outer, 2)
        try:
            inst.call(F_SINGLE_CELL, me, drv, index)
        except Trap:
            # after the 18 uploads single_cell creates and submits an EMPTY command encoder through the device's
            # `dyn Context` (lbm.rs:1511-1514): an indirect call our zeroed Driver cannot serve.  All data is out by then.
            pass
        assert len(got) == 18 and all(len(v) == x * y for v in got.values()) and not inst.called
        return np.stack([np.stack([got[(b, k)] for k in range(9)]) for b in range(2)]).reshape(2, 9, y, x)

    def draw_shape(self, erase, p1, p2, xdim, ydim):
        """LBM::draw_shape(&mut self, &driver, &dyn Shape) (lbm.rs:1337-1343, get_points_vector of merge_shapes.rs
        inlined) on a Curve holding the line p1 -> p2, through a fabricated `dyn Shape` vtable whose get_points is
        served by the host; returns (the u32 array uploaded to draw_points, the word uploaded to draw_num, the
        shape's points).  The function traps at its first wgpu call after the two uploads."""
        got = []
        inst = Instance(self.m, imports=self.stubs,
                        hooks={F_WRITE_BUFFER: lambda i, q, b, ptr, n: got.append(
                            np.frombuffer(i.read(ptr, n), dtype=np.uint32).copy())})
        malloc = self.m.exports["__wbindgen_malloc"][1]
        me, drv, cur, vt = (inst.call(malloc, n, 8) for n in (2048, 1024, 64, 32))
        for a, n in ((me, 2048), (drv, 1024), (cur, 64)):
            inst.mem[a:a + n] = bytes(n)
        struct.pack_into("<QQIIIIIii", inst.mem, cur, 1, 2, 0, 0, 0, self.empty_group, 1, p1[0], p1[1])
        inst.call(F_ERASE_SEGMENT if erase else F_ADD_SEGMENT, cur, p2[0] & 0xFFFFFFFF, p2[1] & 0xFFFFFFFF, xdim, ydim)
        pts = self._read_set(inst, cur)
        struct.pack_into("<I", inst.mem, me + LBM_X, xdim)
        virt = 1 << 20  # a table index the module does not have: served by virtual_table
        struct.pack_into("<IIII", inst.mem, vt, 0, 44, 8, virt)  # {drop, size, align, get_points}
        inst.virtual_table[virt] = lambda i, this: this  # Curve::get_points(&self) -> &self.points (offset 0)
        try:
            inst.call(F_DRAW_SHAPE, me, drv, cur, vt)
        except Trap:
            pass  # queue.submit(None) through the zeroed Driver's `dyn Context`, after both uploads
        assert len(got) == 2 and len(got[1]) == 1 and not inst.called
        return got[0], int(got[1][0]), pts

    def curve_chain(self, ops, xdim, ydim):
        """Curve::new() followed by add_segment / erase_segment calls on the same value (curve.rs:20-48):
        ops = [(erase, x, y), ...]; returns the curve's points (a cell may be present with both flags)"""
        inst, cur = self._instance()
        struct.pack_into("<QQIIIIIii", inst.mem, cur, 1, 2, 0, 0, 0, self.empty_group, 0, 0, 0)  # last_point: None
        for erase, x, y in ops:
            inst.call(F_ERASE_SEGMENT if erase else F_ADD_SEGMENT, cur, x & 0xFFFFFFFF, y & 0xFFFFFFFF, xdim, ydim)
            assert inst.u32(cur + 32) == 1 and (inst.i32(cur + 36), inst.i32(cur + 40)) == (x, y)
        return self._read_set(inst, cur)

    def _mock_wgpu(self):
        """An instance with a fabricated Driver and LBM for running methods of the binary against a MOCK of wgpu's
        `dyn DynContext`: every word of the Driver points at one object that serves both as ArcInner and as vtable,
        whose method slots are served by the host and logged.  Every word of the LBM points into an arena (64 bytes
        per word), so a `&self.field` argument decodes to the field's offset and `&self.vec[i]` to (offset of the
        Vec, i).  Returns (instance, self pointer, driver pointer, command stream) — the stream fills as the code
        runs: ["encoder"], ["write", field, bytes], ["pass", label, pipeline field, [[slot, field, index | None],
        ..], dispatch (field offset, or the number itself if it is not a word of self)], ["submit"], ..."""
        inst = Instance(self.m, imports=self.stubs)
        malloc = self.m.exports["__wbindgen_malloc"][1]
        virt, nslot, rev = 1 << 20, 256, {v: k for k, v in SLOT.items()}
        uni = inst.call(malloc, 4 * nslot, 8)  # as vtable: {drop, size, align = 8, slots..}; as ArcInner: data at +8
        struct.pack_into("<III", inst.mem, uni, virt, 0, 8)
        for k in range(3, nslot):
            struct.pack_into("<I", inst.mem, uni + 4 * k, virt + k)
        anyvt = inst.call(malloc, 16, 4)  # Box<dyn Any> of returned objects: zero-sized, no-op drop
        struct.pack_into("<IIII", inst.mem, anyvt, virt + 250, 0, 1, virt + 251)
        me, drv = inst.call(malloc, 4096, 8), inst.call(malloc, 1024, 8)
        arena = inst.call(malloc, 1024 * 64 + 64, 64)
        inst.mem[arena:arena + 1024 * 64] = bytes(1024 * 64)
        for off in range(0, 4096, 4):
            struct.pack_into("<I", inst.mem, me + off, arena + (off // 4) * 64)
        for off in range(0, 1024, 4):
            struct.pack_into("<I", inst.mem, drv + off, uni)
        stream, cur, ids = [], [None], [100]

        def field(v):
            if me <= v < me + 4096:
                return [v - me, None]
            if not arena <= v < arena + 1024 * 64:
                return [None, v]  # not a word of self: a computed number
            assert (v - arena) % 64 % 24 == 0  # a wgpu::BindGroup is 24 bytes in this build
            return [(v - arena) // 64 * 4, (v - arena) % 64 // 24]

        def method(k):
            def h(i, *a):
                name = rev.get(k)
                if name in ("device_create_command_encoder", "begin_compute_pass", "encoder_finish", "queue_submit"):
                    out = a[0]
                    ids[0] += 1
                    struct.pack_into("<QII", i.mem, out, ids[0], 8, anyvt)  # (ObjectId, Box<dyn Any>)
                    if name == "device_create_command_encoder":
                        assert a[2] == drv + DRIVER_DEVICE
                        stream.append(["encoder"])
                    elif name == "begin_compute_pass":
                        p, ln = i.u32(a[-1]), i.u32(a[-1] + 4)  # ComputePassDescriptor { label: Option<&str> }
                        cur[0] = ["pass", i.read(p, ln).decode() if p else None, None, [], None]
                    elif name == "queue_submit":
                        assert a[2] == drv + DRIVER_QUEUE
                        stream.append(["submit"])
                    return None
                if name == "queue_write_buffer":
                    assert a[1] == drv + DRIVER_QUEUE and a[7] == 0  # offset 0
                    stream.append(["write", a[4] - me - 48, i.read(a[8], a[9]).hex()])  # &buffer.id is 48 bytes in
                elif name == "set_pipeline":
                    cur[0][2] = field(a[4])[0]
                elif name == "set_bind_group":
                    assert a[9] == 0  # no dynamic offsets
                    cur[0][3].append([a[4]] + field(a[5]))
                elif name == "dispatch_workgroups":
                    assert a[5] == 1 and a[6] == 1
                    f = field(a[4])  # the VALUE of a self word = its arena slot: which field was read
                    cur[0][4] = f[0] if f[0] is not None else f[1]
                elif name == "end_compute_pass":
                    stream.append(cur[0])
                    cur[0] = None
                elif k in (250, 251):
                    pass
                else:
                    raise Trap(f"unmocked DynContext slot {k}")
                return 0
            return h

        for k in range(nslot):
            inst.virtual_table[virt + k] = method(k)
        return inst, me, drv, stream

    def iterate_trace(self, n, compute_step=0, stat=0, cmap=2):
        """LBM::iterate(&mut self, &driver, n) (lbm.rs:1065-1074: n x (collide; stream; compute_step += 1), then
        calculate_summary + color_map) against the mock wgpu context; returns the command stream.  The run ends
        when iterate reaches render() (the surface is not mocked)."""
        inst, me, drv, stream = self._mock_wgpu()
        struct.pack_into("<II", inst.mem, me + LBM_X, 64, 32)
        struct.pack_into("<I", inst.mem, me + LBM_COMPUTE_STEP, compute_step)
        struct.pack_into("<I", inst.mem, me + LBM_STAT_CMAP, stat | (cmap << 8))
        try:
            inst.call(F_ITERATE, me, drv, n)
        except Trap:
            pass  # render(): surface_get_current_texture is not mocked
        assert inst.u32(me + LBM_COMPUTE_STEP) == compute_step + n and not inst.called
        return stream

    def draw_trace(self, p1, p2, xdim, ydim):
        """LBM::draw_shape(&mut self, &driver, &Curve{line p1 -> p2}) against the mock wgpu context, to completion:
        the two uploads, the empty submit, and the barrier_draw pass with its dispatch size (lbm.rs:1341-1356)"""
        inst, me, drv, stream = self._mock_wgpu()
        malloc = self.m.exports["__wbindgen_malloc"][1]
        cur, vt = inst.call(malloc, 64, 8), inst.call(malloc, 32, 4)
        struct.pack_into("<QQIIIIIii", inst.mem, cur, 1, 2, 0, 0, 0, self.empty_group, 1, p1[0], p1[1])
        inst.call(F_ADD_SEGMENT, cur, p2[0] & 0xFFFFFFFF, p2[1] & 0xFFFFFFFF, xdim, ydim)
        npoints = len(self._read_set(inst, cur))
        struct.pack_into("<I", inst.mem, me + LBM_X, xdim)
        struct.pack_into("<IIII", inst.mem, vt, 0, 44, 8, (1 << 20) + 200)
        inst.virtual_table[(1 << 20) + 200] = lambda i, this: this  # Curve::get_points(&self) -> &self.points
        inst.call(F_DRAW_SHAPE, me, drv, cur, vt)  # returns normally
        assert not inst.called
        return npoints, stream

    def line_new(self, p1, p2, xdim, ydim):
        inst, ret = self._instance()
        inst.call(F_LINE_NEW, ret, *(v & 0xFFFFFFFF for v in (p1[0], p1[1], p2[0], p2[1], xdim, ydim)))
        return self._read_set(inst, ret)

    def curve_segment(self, erase, p1, p2, xdim, ydim):
        """Curve { points: empty, last_point: Some(p1) }.erase_segment(p2, ..) / .add_segment(p2, ..)"""
        inst, cur = self._instance()
        struct.pack_into("<QQIIIIIii", inst.mem, cur, 1, 2, 0, 0, 0, self.empty_group, 1, p1[0], p1[1])
        inst.call(F_ERASE_SEGMENT if erase else F_ADD_SEGMENT, cur, p2[0] & 0xFFFFFFFF, p2[1] & 0xFFFFFFFF, xdim, ydim)
        assert (inst.i32(cur + 36), inst.i32(cur + 40)) == tuple(p2)  # last_point = Some(next)
        return self._read_set(inst, cur)


class _Done(Exception):
    pass


def locate_presets(m):
    """(function index, {name: (start pc, number of draw_shape calls)}, lbm local, driver local) of the three preset
    arms, found by structure: the closure is the function with >= 30 calls of Line::new; its calls of draw_shape
    split them into welcome (28 lines, one Blob drawn), curl (1 line, 1 draw) and chaos (4 lines, 4 draws); an arm
    starts where its shadow-stack frame is opened (`global.get 0` right after the `end` of the previous arm)."""
    cands = []
    for k in range(len(m.bodies)):
        f = k + m.n_imports
        code = m.decode(f)[1]
        n = sum(1 for i in code if i[0] == 0x10 and i[1] == F_LINE_NEW)
        if n >= 30:
            cands.append(f)
    assert len(cands) == 1, cands
    f = cands[0]
    code = m.decode(f)[1]
    lines = [pc for pc, i in enumerate(code) if i[0] == 0x10 and i[1] == F_LINE_NEW]
    draws = [pc for pc, i in enumerate(code) if i[0] == 0x10 and i[1] == F_DRAW_SHAPE]
    # arms in code order: 28 lines + 1 draw, 1 line + 1 draw, then (line, draw) x 4
    assert all(a < draws[0] for a in lines[:28]) and lines[28] > draws[0]
    assert draws[0] < lines[28] < draws[1] < lines[29] < draws[2] < lines[30] < draws[3] < lines[31] < draws[4] < lines[32] < draws[5]

    def arm_start(first_line_pc):
        pc = first_line_pc
        while not (code[pc][0] == 0x23 and code[pc][1] == 0 and code[pc - 1][0] == 0x0B):
            pc -= 1
            assert first_line_pc - pc < 120
        return pc

    arms = {"welcome": (arm_start(lines[0]), 1), "curl": (arm_start(lines[28]), 1), "chaos": (arm_start(lines[29]), 4)}
    # the locals: draw_shape(&mut lbm, &driver, &shape, vtable) — the first two arguments of the curl arm's call
    pc = draws[1]
    j = pc - 3
    while not (code[j][0] == 0x20 and code[j + 1][0] == 0x20 and code[j + 2][0] == 0x20):  # lbm, driver, shape base
        j -= 1
        assert pc - j < 12
    lbm_local, drv_local = code[j][1], code[j + 1][1]
    s = arms["curl"][0]
    assert any(i[0] == 0x28 and i[1] == LBM_X for i in code[s:s + 12]) and [code[s + 6][0], code[s + 6][1]] == [0x20, lbm_local]
    return f, arms, lbm_local, drv_local


def preset_trace(ref, name, xdim, ydim):
    """Run one preset arm of the event-loop closure: returns (the (x1, y1, x2, y2, xdim, ydim) of every Line::new
    call in order, the sorted (x, y, on) point list of every draw_shape call in order)."""
    m = ref.m
    f, arms, lbm_local, drv_local = locate_presets(m)
    start, ndraws = arms[name]
    lines, draws = [], []
    s32 = lambda v: v - (1 << 32) if v >= 1 << 31 else v  # noqa: E731

    def line_new(i, out, *a):
        lines.append(tuple(s32(v) for v in a))
        del i.hooks[F_LINE_NEW]
        try:
            i._invoke(F_LINE_NEW, [out, *a])
        finally:
            i.hooks[F_LINE_NEW] = line_new

    def draw_shape(i, me_, drv_, shape, vt):
        assert (me_, drv_) == (me, drv)
        get_points = m.table[i.u32(vt + 12)]  # {drop, size, align, get_points}
        draws.append(Reference._read_set(i, i._invoke(get_points, [shape])[0]))
        if len(draws) == ndraws:
            raise _Done

    inst = Instance(m, imports=ref.stubs, hooks={F_LINE_NEW: line_new, F_DRAW_SHAPE: draw_shape})
    malloc = m.exports["__wbindgen_malloc"][1]
    me, drv = inst.call(malloc, 2048, 8), inst.call(malloc, 1024, 8)
    for a, n in ((me, 2048), (drv, 1024)):
        inst.mem[a:a + n] = bytes(n)
    struct.pack_into("<II", inst.mem, me + LBM_X, xdim, ydim)
    try:
        inst.run_fragment(f, start, {lbm_local: me, drv_local: drv})
    except _Done:
        pass
    assert len(draws) == ndraws and not inst.called
    return lines, draws


def locate_lbm_new(m):
    """`LBM::new` (lbm.rs:726-1049) is inlined into the async start-up closure (`run_wasm`, the one other function
    besides the event loop that calls set_equil with the constant 0.1 in front).  Its wgpu object creations are NOT
    inlined: nine calls of Device::create_bind_group_layout, then set_equil, create_buffer_init (the callee with
    several context calls), create_bind_group, create_pipeline_layout, create_shader_module, create_compute_pipeline,
    create_render_pipeline — told apart by how often the closure (or its helpers) call them.  Returns the closure,
    the pc where LBM::new starts (x and y are read from the future's state), the creation functions and the Drop
    impls of wgpu objects (one-argument functions that call the context once)."""
    start_f = [k + m.n_imports for k in range(len(m.bodies))
               if any(i[0] == 0x10 and i[1] == F_SET_EQUIL and j and m.decode(k + m.n_imports)[1][j - 3][0] == 0x43
                      for j, i in enumerate(m.decode(k + m.n_imports)[1]))]
    start_f = [f for f in start_f if len(m.decode(f)[1]) > 5000 and sum(1 for i in m.decode(f)[1] if i[0] == 0x11) > 20]
    assert len(start_f) == 1, start_f
    f = start_f[0]
    code = m.decode(f)[1]
    eq = next(pc for pc, i in enumerate(code) if i[0] == 0x10 and i[1] == F_SET_EQUIL)

    def one_ctx_call(fn):
        c = m.decode(fn)[1]
        return fn >= m.n_imports and m.type_of(fn) == ([I32] * 3, []) and sum(1 for i in c if i[0] == 0x11) == 1

    import collections
    before = collections.Counter(i[1] for i in code[:eq] if i[0] == 0x10 and one_ctx_call(i[1]))
    (bgl, nb), = [(k, v) for k, v in before.items() if v == 9]
    # LBM::new ends with the render pipeline, ~3300 instructions after set_equil (the closure goes on to build the
    # event loop, which creates more objects)
    after = collections.Counter(i[1] for i in code[eq:eq + 3300] if i[0] == 0x10 and one_ctx_call(i[1]))
    by_count = {v: k for k, v in after.items()}
    assert sorted(after.values()) == [1, 6, 8, 16, 17], after  # render pipeline, layouts, bind groups, pipelines, shaders
    buf_init = [k for k in {i[1] for i in code[eq:eq + 800] if i[0] == 0x10 and i[1] >= m.n_imports}
                if sum(1 for j in m.decode(k)[1] if j[0] == 0x11) >= 5]
    assert len(buf_init) == 1
    drops = [k + m.n_imports for k in range(len(m.bodies))
             if m.type_of(k + m.n_imports) == ([I32], []) and len(m.decode(k + m.n_imports)[1]) < 60
             and sum(1 for i in m.decode(k + m.n_imports)[1] if i[0] == 0x11) == 1]
    first = next(pc for pc, i in enumerate(code) if i[0] == 0x10 and i[1] == bgl)
    pc = first
    while not (code[pc][0] == 0x23 and code[pc][1] == 0):  # the frame of LBM::new is opened with `global.get 0`
        pc -= 1
        assert first - pc < 80
    start = pc - 6  # local.get 0; i32.load w; local.set; local.get 0; i32.load h; local.set
    assert [code[start][0], code[start + 1][0], code[start + 3][0], code[start + 4][0]] == [0x20, 0x28, 0x20, 0x28]
    assert code[start + 4][1] == code[start + 1][1] + 4
    # the Driver sits at a fixed offset of the closure's own frame: `&driver.device` = frame + k + DRIVER_DEVICE
    return {"f": f, "start": start, "state_x": code[start + 1][1], "state_local": code[start][1],
            "bgl": bgl, "bg": by_count[8], "pl": by_count[6], "shader": by_count[17], "cpipe": by_count[16],
            "rpipe": by_count[1], "buffer_init": buf_init[0], "drops": drops}


def bindgroup_trace(ref, x=12, y=7):
    """Run LBM::new inside the start-up closure of the binary (fragment execution from where it reads x and y) with
    the wgpu object creations served by the host: every returned object is filled with a pointer to a tag (which
    doubles as a harmless ArcInner / Box vtable), so that it can be recognised wherever the code stores or passes it.
    Returns what the binary itself did: the buffers it created (order, contents, usage), the entries of every bind
    group (binding -> buffer), where each buffer of `data_buffers[b][k]` and each bind group ends up in the assembled
    `LBM` value (field offsets — the very offsets `LBM::iterate` passes to set_bind_group, see iterate_trace)."""
    m = ref.m
    loc = locate_lbm_new(m)
    inst = Instance(m, imports=ref.stubs, max_steps=400_000_000)
    malloc = m.exports["__wbindgen_malloc"][1]
    virt, nslot, RC = 1 << 20, 256, 0x40000000
    uni = inst.call(malloc, 4 * nslot, 8)  # as in _mock_wgpu: the context's ArcInner and vtable in one object
    struct.pack_into("<III", inst.mem, uni, virt, 0, 8)
    for k in range(3, nslot):
        struct.pack_into("<I", inst.mem, uni + 4 * k, virt + k)
    anyvt = inst.call(malloc, 16, 4)
    struct.pack_into("<IIII", inst.mem, anyvt, virt + 250, 0, 1, virt + 251)
    state, frame, arena = inst.call(malloc, 8192, 8), inst.call(malloc, 8192, 8), inst.call(malloc, 64 * 512, 8)
    for a, n in ((state, 8192), (frame, 8192), (arena, 64 * 512)):
        inst.mem[a:a + n] = bytes(n)
    for off in range(16, 2048, 4):
        struct.pack_into("<I", inst.mem, frame + off, uni)  # every word of the Driver -> the mock context
    struct.pack_into("<II", inst.mem, state + loc["state_x"], x, y)
    tags, buffers, groups, layouts, created = {}, [], [], [], []

    def new_tag(kind, n):
        t = arena + 64 * len(tags)
        tags[t] = (kind, n)
        inst.write_u32(t, RC)  # a strong count that never reaches zero / a vtable whose drop is the no-op below
        return t

    def fill(ret, nbytes, t):
        for o in range(0, nbytes, 4):
            inst.write_u32(ret + o, t)

    def buffer_of(ptr):
        """which buffer a `&wgpu::Buffer` points at: a tag in its first word (create_buffer_init, served whole) or
        the ObjectId the mock context handed out (Device::create_buffer is inlined; &buffer.id is 48 bytes in)"""
        t = inst.u32(ptr)
        if t in tags and tags[t][0] == "buffer":
            return tags[t][1]
        oid = inst.u32(ptr + 48)
        return next(b["index"] for b in buffers if b.get("object_id") == oid)

    def create_buffer_init(i, ret, dev, desc):  # BufferInitDescriptor { label, contents: &[u8], usage }
        lp, ll, cp, cl, usage = (i.u32(desc + 4 * k) for k in range(5))
        data = i.read(cp, cl)
        w = np.frombuffer(data[:cl // 4 * 4], np.uint32)
        buffers.append({"index": len(buffers), "how": "create_buffer_init", "label": i.read(lp, ll).decode() if lp else None,
                        "bytes": cl, "usage": usage, "first_word": int(w[0]) if len(w) else None,
                        "uniform_contents": bool(len(w) and (w == w[0]).all()),
                        "sha256": hashlib.sha256(data).hexdigest()})
        fill(ret, SIZEOF_BUFFER, new_tag("buffer", len(buffers) - 1))

    def create_bind_group(i, ret, dev, desc):  # BindGroupDescriptor { label, entries: &[BindGroupEntry], layout }
        ep, en, lay = i.u32(desc + 8), i.u32(desc + 12), i.u32(desc + 16)
        entries = []
        for k in range(en):  # BindGroupEntry (40 bytes): the &Buffer of BufferBinding at +24, `binding` at +32
            e = ep + 40 * k
            entries.append([i.u32(e + 32), buffer_of(i.u32(e + 24))])
        groups.append({"index": len(groups), "layout": tags[i.u32(lay)][1], "entries": entries})
        fill(ret, 24, new_tag("bind_group", len(groups) - 1))

    def creator(kind):
        def h(i, ret, dev, desc):
            n = sum(1 for c in created if c == kind)
            created.append(kind)
            fill(ret, 24, new_tag(kind, n))
        return h

    class LbmAssembled(Exception):
        pass

    oid, create_slot, other_slots = [1000], [None], []

    def method(k):
        def h(i, *a):
            if k in (250, 251):
                return None
            if len(groups) == 16 and i.u32(frame + lbm_base + LBM_X) == x and i.u32(frame + lbm_base + LBM_Y) == y:
                raise LbmAssembled  # the first context call after LBM::new's value has been moved into place
            if create_slot[0] is None:
                create_slot[0] = k  # the first context call made inline: Device::create_buffer (the colour buffer)
            oid[0] += 1
            if k != create_slot[0]:
                # what follows LBM::new in the closure (surface configuration, the first frame): answered like a
                # creation — (ObjectId, Box<dyn Any>) through the out pointer — until LBM's value is in place
                other_slots.append(k)
                struct.pack_into("<QII", i.mem, a[0], oid[0], 8, anyvt)
                return None
            # Device::create_buffer(&BufferDescriptor { label, size: u64, usage, mapped_at_creation }), inlined
            d = a[-1]
            buffers.append({"index": len(buffers), "how": f"create_buffer (context slot {k})", "object_id": oid[0],
                            "descriptor_words": [i.u32(d + 4 * q) for q in range(6)]})
            struct.pack_into("<QII", i.mem, a[0], oid[0], 8, anyvt)
            return None
        return h

    for k in range(nslot):
        inst.virtual_table[virt + k] = method(k)
    for d in range(-256, 256):
        inst.virtual_table[RC + d] = lambda i, *a: None
    inst.hooks.update({loc["bgl"]: creator("bind_group_layout"), loc["pl"]: creator("pipeline_layout"),
                       loc["shader"]: creator("shader_module"), loc["cpipe"]: creator("compute_pipeline"),
                       loc["rpipe"]: creator("render_pipeline"), loc["buffer_init"]: create_buffer_init,
                       loc["bg"]: create_bind_group})
    for f in loc["drops"]:
        inst.hooks[f] = lambda i, a: None
    lbm_base = 168  # the `lbm` local of the closure's frame (checked below through x and y)
    try:
        inst.run_fragment(loc["f"], loc["start"], {loc["state_local"]: state, 4: frame})
        raise AssertionError("the closure returned before LBM::new's value was seen")
    except (LbmAssembled, Trap):
        pass  # (a Trap: the first frame's render path, which the mock does not serve; the value is in place by then)
    base = frame + lbm_base
    assert (inst.u32(base + LBM_X), inst.u32(base + LBM_Y)) == (x, y) and not inst.called

    def tag_at(addr):
        t = inst.u32(addr)
        return list(tags[t]) if t in tags else None

    # the assembled LBM value: which tagged object sits at which field offset (24-byte objects; Vec {cap, ptr, len})
    fields = {}
    for off in range(0, 1400, 4):
        t = tag_at(base + off)
        if t and all(tag_at(base + off + 4 * q) == t for q in range(6)) and (off < 24 or tag_at(base + off - 4) != t or
                                                                              (off - min(o for o in fields or [off])) % 24 == 0):
            if not any(o < off < o + (SIZEOF_BUFFER if v[0] == "buffer" else 24) for o, v in fields.items()):
                fields[off] = t
    vecs = {}
    for off in range(0, 1400, 4):  # Vec<BindGroup> / Vec<Vec<Buffer>>: ptr word at off, len right behind
        ptr, ln = inst.u32(base + off), inst.u32(base + off + 4)
        if ln == 2 and 0x1000 < ptr < len(inst.mem) - 48 and not arena <= ptr < arena + 64 * 512:
            a, b = tag_at(ptr), tag_at(ptr + 24)
            if a and b and a[0] == b[0] == "bind_group":
                vecs[off] = [a[1], b[1]]
    outer_ptr, outer_len = inst.u32(base + LBM_BUFFERS_PTR), inst.u32(base + LBM_BUFFERS_LEN)
    assert outer_len == 2
    data_buffers = []
    for b in range(2):
        p, ln = inst.u32(outer_ptr + 12 * b + 4), inst.u32(outer_ptr + 12 * b + 8)
        assert ln == 9
        data_buffers.append([tag_at(p + SIZEOF_BUFFER * k)[1] for k in range(9)])
    return {"x": x, "y": y, "buffers": buffers, "bind_groups": groups, "created": created,
            "lbm_fields": {str(k): v for k, v in sorted(fields.items())},
            "lbm_bind_group_vecs": {str(k): v for k, v in sorted(vecs.items())}, "data_buffers": data_buffers}


def presets_main(ref):
    out = {"wasm_sha256": np.bytes_(ref.sha256), "sizes": np.array(PRESET_SIZES, np.int64)}
    for x, y in PRESET_SIZES:
        for name in ("curl", "chaos", "welcome"):
            lines, draws = preset_trace(ref, name, x, y)
            out[f"{name}/{x}x{y}/lines"] = np.array(lines, np.int64).reshape(-1, 6)
            for k, d in enumerate(draws):
                out[f"{name}/{x}x{y}/draw{k}"] = np.array(d, np.int32).reshape(-1, 3)
            print(f"{name}_barrier on {x}x{y}: {len(lines)} lines, draws of {[len(d) for d in draws]} cells", flush=True)
    path = os.path.join(HERE, "wasm_presets.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
    import json
    bg = bindgroup_trace(ref)
    bg["wasm_sha256"] = ref.sha256
    path = os.path.join(HERE, "wasm_bindgroups.json")
    with open(path, "w") as f:
        json.dump(bg, f, indent=1)
    print(path, os.path.getsize(path), "bytes:", len(bg["buffers"]), "buffers,", len(bg["bind_groups"]), "bind groups")


def main():
    ref = Reference()
    if "presets" in sys.argv[1:]:
        presets_main(ref)
        return
    presets_main(ref)
    out = {"wasm_sha256": np.bytes_(ref.sha256), "inflows": np.array(INFLOWS, np.float32)}
    out["set_equil"] = np.stack([ref.set_equil(u) for u in INFLOWS])
    for x, y in SINGLE_CELL_SIZES:
        for index in range(10):
            out[f"single_cell/{x}x{y}/{index}"] = ref.single_cell(index, x, y)
        print(f"single_cell on {x}x{y}: 10 presets", flush=True)
    out["single_cell_sizes"] = np.array(SINGLE_CELL_SIZES, np.int64)
    cases = line_cases()
    n_draw = 0
    for i, (a, b, xd, yd) in enumerate(cases):
        if xd == 64 and i % 4 == 0 and min(a + b) >= 0 and max(a[0], b[0]) < xd and max(a[1], b[1]) < yd:
            for erase in (0, 1):
                pairs, count, pts = ref.draw_shape(bool(erase), a, b, xd, yd)
                out[f"draw/{i}/{erase}/pairs"] = pairs.reshape(-1, 2)
                out[f"draw/{i}/{erase}/count"] = np.int64(count)
                out[f"draw/{i}/{erase}/points"] = np.array(pts, np.int32).reshape(-1, 3)
            n_draw += 1
    print(f"draw_shape: {n_draw} lines, drawn and erased", flush=True)
    out["line_cases"] = np.array([[a[0], a[1], b[0], b[1], xd, yd] for a, b, xd, yd in cases], np.int64)
    for i, (a, b, xd, yd) in enumerate(cases):
        pts = ref.line_new(a, b, xd, yd)
        if pts is None:
            out[f"line/{i}/new"] = np.zeros((0, 3), np.int32)  # Err(..): a valid line has at least one cell
            continue
        out[f"line/{i}/new"] = np.array(pts, np.int32).reshape(-1, 3)
        seg = ref.curve_segment(False, a, b, xd, yd)
        assert seg == pts, "Curve::add_segment must give Line::new's points"
        if i % 3 == 0 or xd == 64:  # the 30-wide eraser is slow to interpret: a subset of the cases
            out[f"line/{i}/erased"] = np.array(ref.curve_segment(True, a, b, xd, yd), np.int32).reshape(-1, 3)
        print(f"case {i}: {a} -> {b} on {xd}x{yd}: {len(pts)} cells", flush=True)
    for j, (ops, xd, yd) in enumerate(curve_chains()):
        out[f"curve/{j}/ops"] = np.array(ops, np.int64)
        out[f"curve/{j}/dims"] = np.array([xd, yd], np.int64)
        out[f"curve/{j}/points"] = np.array(ref.curve_chain(ops, xd, yd), np.int32).reshape(-1, 3)
        print(f"curve chain {j}: {len(ops)} segments -> {len(out[f'curve/{j}/points'])} points", flush=True)
    import json
    out["iterate_trace/steps3_from0"] = np.bytes_(json.dumps(ref.iterate_trace(3, compute_step=0)))
    out["iterate_trace/steps2_from7"] = np.bytes_(json.dumps(ref.iterate_trace(2, compute_step=7)))
    for stat in range(5):
        for cmap in range(3):
            out[f"iterate_trace/frame_only/{stat}/{cmap}"] = np.bytes_(json.dumps(ref.iterate_trace(0, stat=stat, cmap=cmap)))
    npoints, stream = ref.draw_trace((3, 4), (20, 11), 64, 32)
    out["draw_trace/npoints"] = np.int64(npoints)
    out["draw_trace/stream"] = np.bytes_(json.dumps(stream))
    print("iterate, draw_shape: command streams recorded", flush=True)
    path = os.path.join(HERE, "wasm_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
