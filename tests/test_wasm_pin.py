"""Parity pinned to OUTPUTS OF THE REFERENCE ITSELF for its host-side rows.

tests/golden/wasm_golden.npz holds the results of executing functions of the reference's own shipped binary
(lbm-wgpu/pkg/lbm_wgpu_bg.wasm, run in oracle/wasm_mini.py by tests/golden/make_wasm_golden.py):
  * `set_equil` (lbm.rs:611-643) -> the nine initial populations for eight inflow speeds   
  * `LBM::single_cell` (lbm.rs:1482-1515) -> the 18 arrays it uploads, for every preset index on four lattice sizes (row a-8)
  * `LBM::draw_shape` (lbm.rs:1337-1343 + merge_shapes.rs:12-22) -> the u32 pairs and the count word it uploads for
    barrier_draw.wgsl, for drawn and erased lines: the wire format blbm_draw_points takes                 (row a-7)
  * `LBM::iterate` (lbm.rs:1065-1074, everything below it inlined) against a mock of wgpu's `dyn DynContext` -> the
    complete command stream: encoders, labelled compute passes, pipeline and bind-group fields, alternation with
    compute_step % 2, summary and colour-map passes — the host-side dispatch order                  (rows a-3 .. a-6)
  * `Curve::add_segment` / `erase_segment` chains (curve.rs:20-48) -> the points of whole strokes, against the C++ mirror
  * `Line::new` / `Line::new_erased` (barrier_shapes/line.rs:22-87, with the un-vendored line_drawing 1.0.0
    Bresenham as compiled in) -> the cells of 99 thick lines on two lattice sizes           (row N2)
Here the oracle's restatements, the product's host-side rasteriser (libblbm.so: blbm_rasterize_line is pure host
code and runs without a GPU) and, on the GPU, the initial state and the painted mask of the CUDA path are compared
with them.  Where /root/reference exists the binary is re-run live."""
import os

import numpy as np
import pytest

from oracle import barrier_shapes
from oracle.lbm_oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "wasm_golden.npz")
WASM = "/root/reference/lbm-wgpu/pkg/lbm_wgpu_bg.wasm"
needs_reference = pytest.mark.skipif(not os.path.exists(WASM), reason="/root/reference is only present in the build container")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def cases(golden):
    for i, (x1, y1, x2, y2, xd, yd) in enumerate(golden["line_cases"].tolist()):
        yield i, (x1, y1), (x2, y2), xd, yd


def cells(a):
    return sorted(map(tuple, np.asarray(a)[:, :2].tolist()))


def test_oracle_initial_populations_equal_the_reference_binarys_set_equil(golden):
    """the C oracle, the numpy restatement and the WGSL driver's host-side set_equil against the compiled one"""
    from oracle.lbm_numpy import NumpyLBM
    from oracle.wgsl_interp import WgslLBM
    for ux, want in zip(golden["inflows"], golden["set_equil"]):
        o = Oracle(1.0, 6, 5, inflow_ux=float(ux))
        got = np.array([o.population(0, k).ravel()[7] for k in range(9)], np.float32)
        assert (bits(got) == bits(want)).all(), (ux, got, want)
        for b in (0, 1):
            for k in range(9):
                assert (bits(o.population(b, k)) == bits(want[k])).all()
        o.close()
        w = np.array(WgslLBM.set_equil(np.float32(ux), np.float32(0), np.float32(1)), np.float32)
        assert (bits(w) == bits(want)).all()
        n = NumpyLBM(1.0, 6, 5, inflow_ux=float(ux))
        assert all((bits(n.population(0, k)) == bits(want[k])).all() for k in range(9))


def single_cell_cases(golden):
    for x, y in golden["single_cell_sizes"].tolist():
        for index in range(10):
            yield x, y, index, golden[f"single_cell/{x}x{y}/{index}"]  # [buffer][population][y][x]


def test_oracle_single_cell_equals_what_the_reference_binary_uploads(golden):
    """LBM::single_cell as compiled: which population holds the 4.0 packet at which cell, everything else
    set_equil(0, 0, 1); the reference uploads the rest population to both data_buffers[b][4], of which only
    buffer 0's is ever bound (lbm.rs:775-778) — the oracle keeps that one"""
    n = 0
    for x, y, index, want in single_cell_cases(golden):
        assert (bits(want[0]) == bits(want[1])).all()
        o = Oracle(1.0, x, y)
        o.iterate(3)
        o.single_cell(index)
        for b in (0, 1):
            for k in range(9):
                assert (bits(np.asarray(o.population(b, k)).reshape(y, x)) == bits(want[b, k])).all(), (x, y, index, b, k)
        assert o.get_compute_num() == 0
        o.close()
        n += 1
    assert n == 40


def test_host_rasteriser_equals_the_reference_binarys_lines(golden):
    """libblbm.so's blbm_rasterize_line (the product's host code in front of draw_points) and the oracle's
    restatement against Line::new / Line::new_erased as compiled into the reference's binary"""
    from lbm_b200.lbm import rasterize_line
    n_new = n_erased = n_err = 0
    for i, p1, p2, xd, yd in cases(golden):
        want = golden[f"line/{i}/new"]
        ours = rasterize_line(p1, p2, xd, yd, erase=False)
        ora = barrier_shapes.line_points(p1, p2, xd, yd, erase=False)
        if len(want) == 0:  # the reference returned Err(..)
            assert ours is None and ora is None
            n_err += 1
            continue
        assert (want[:, 2] == 1).all()
        assert cells(ours) == cells(want), f"Line::new {p1}->{p2} on {xd}x{yd}"
        assert sorted((x, y) for x, y, *_ in ora) == cells(want)
        n_new += 1
        key = f"line/{i}/erased"
        if key in golden.files:
            want = golden[key]
            assert (want[:, 2] == 0).all()
            assert cells(rasterize_line(p1, p2, xd, yd, erase=True)) == cells(want), f"Line::new_erased {p1}->{p2}"
            assert sorted((x, y) for x, y, *_ in barrier_shapes.line_points(p1, p2, xd, yd, erase=True)) == cells(want)
            n_erased += 1
    assert n_new >= 90 and n_erased >= 50 and n_err == 3


def draw_cases(golden):
    for key in sorted(k for k in golden.files if k.startswith("draw/") and k.endswith("/pairs")):
        _, i, erase, _ = key.split("/")
        x1, y1, x2, y2, xd, yd = golden["line_cases"][int(i)].tolist()
        yield (x1, y1), (x2, y2), xd, yd, int(erase), golden[key], int(golden[f"draw/{i}/{erase}/count"]), \
            golden[f"draw/{i}/{erase}/points"]


def test_paint_wire_format_equals_what_the_reference_binary_uploads(golden):
    """draw_shape as compiled: the array handed to barrier_draw.wgsl is [x + y*W, 1|0] per point of the shape (hash
    order), and the count word is 2n - 1 (`points.len() as u32 - 1` on the flattened vector, lbm.rs:1343).  The
    product's host-side flattening (lbm_b200.lbm.points_vector, used by LBM.draw_shape) and the host rasteriser
    produce the same set of pairs."""
    from lbm_b200.lbm import points_vector, rasterize_line
    n = 0
    for p1, p2, xd, yd, erase, pairs, count, pts in draw_cases(golden):
        want = sorted(map(tuple, pairs.tolist()))
        assert count == 2 * len(pairs) - 1 and len(pairs) == len(pts)
        assert sorted(map(tuple, points_vector([(x, y, bool(f)) for x, y, f in pts.tolist()], xd).tolist())) == want
        ours = rasterize_line(p1, p2, xd, yd, erase=bool(erase))
        assert sorted((x + y * xd, 0 if erase else 1) for x, y in ours.tolist()) == want
        n += 1
    assert n >= 16


def test_cpp_curve_equals_the_reference_binarys_curve(golden, tmp_path):
    """include/blbm.hpp's Curve (add_segment / erase_segment chains, curve.rs:28-48) compiled and run on the host
    against the same strokes executed in the reference's binary; a cell may carry both flags, as in the reference"""
    import subprocess
    root = os.path.dirname(HERE)
    libdir = os.path.join(root, "lbm_b200")
    exe = str(tmp_path / "curve_chain")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(root, "include"),
                        os.path.join(HERE, "cpp", "curve_chain.cpp"), "-o", exe, "-L", libdir, "-lblbm",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n = 0
    while f"curve/{n}/ops" in golden.files:
        ops, (xd, yd), want = golden[f"curve/{n}/ops"].tolist(), golden[f"curve/{n}/dims"].tolist(), golden[f"curve/{n}/points"]
        r = subprocess.run([exe, str(xd), str(yd)] + [str(v) for op in ops for v in op], capture_output=True, text=True,
                           timeout=60)
        assert r.returncode == 0, r.stdout + r.stderr
        got = sorted(tuple(int(v) for v in line.split()) for line in r.stdout.splitlines())
        assert got == sorted(map(tuple, want.tolist())), f"curve chain {n}: {ops}"
        n += 1
    assert n >= 9


# the labels the reference gives its compute passes (lbm.rs:1174-1252) <-> the shader each pipeline is built from
PASS_OF_LABEL = {"Precollision-corner": "pre_corner", "Precollision-cardinal": "pre_cardinal",
                 "Collision-corner": "col_corner", "Collision-cardinal": "col_cardinal", "Stream_e_w": "e_w",
                 "Stream_n_s": "n_s", "Stream_nw_se": "nw_se", "Stream_ne_sw": "ne_sw"}


def recorded_driver(n, compute_step, stat, cmap):
    """the oracle-side host driver (oracle/wgsl_interp.py: WgslLBM — the dispatch order every golden vector of
    tests/golden/wgsl_*.npz was produced with) run WITHOUT shaders, recording per dispatch what it binds to each
    bind-group slot: [(shader name, {group: identity of what is bound there})]"""
    from oracle.wgsl_interp import REST, WgslLBM

    class Recorder(WgslLBM):
        def _load_shaders(self, root):
            self.sh = {}

        def identity(self, group_bindings):
            ids = set()
            for b in group_bindings:
                hit = None
                for buf in (0, 1):
                    for k in range(9):
                        if b is self.data[buf][k]:
                            hit = ("rest",) if k == REST else ("population", buf, k)
                for nm in ("ux", "uy", "rho", "output", "barrier", "colors"):
                    if b is getattr(self, nm):
                        hit = (nm,)
                if hit is None:
                    hit = ("dims",) if isinstance(b, dict) else ("size",) if isinstance(b, np.uint32) else ("omega",)
                ids.add(hit)
            return frozenset(ids)

        def _run(self, name, bindings):
            groups = {}
            for (g, _), b in sorted(bindings.items()):
                groups.setdefault(g, []).append(b)
            self.log.append((name, {g: self.identity(v) for g, v in groups.items()}))

        def color_map(self, name=None):
            self._run(name or self.color_map_name, {(0, 0): self.colors, (1, 0): self.output, (2, 0): self.barrier,
                                                    (3, 0): np.uint32(self.n)})

    sim = Recorder(1.0, 8, 4)
    sim.log = []
    sim.compute_step = compute_step
    sim.set_summary(Recorder.STATS[stat])
    sim.iterate(n)
    sim.color_map(Recorder.CMAPS[cmap])
    return sim.log


def check_command_stream(stream, n, compute_step, stat, cmap):
    """the reference binary's command stream against the recorded driver: same passes in the same order, and ONE
    consistent correspondence between the binary's bind-group fields (with their compute_step % 2 element) and what
    the driver binds — in both directions"""
    log = recorded_driver(n, compute_step, stat, cmap)
    passes = [r for r in stream if r[0] == "pass"]
    assert len(passes) == len(log) == 8 * n + 2
    # shape of the stream: per step two encoders (4 collide passes, 4 stream passes), then one for summary + colours
    kinds = [r[0] for r in stream]
    assert kinds[:(6 + 6) * n] == (["encoder"] + ["pass"] * 4 + ["submit"]) * (2 * n)
    assert kinds[12 * n:12 * n + 4] == ["encoder", "pass", "pass", "submit"]
    field_to_identity, identity_to_field, pipeline_of = {}, {}, {}
    dispatch_fields = set()
    for (_, label, pipeline, groups, dispatch), (name, bound) in zip(passes, log):
        if label is not None:
            assert PASS_OF_LABEL[label] == name, (label, name)
        assert pipeline_of.setdefault(name, pipeline) == pipeline  # one pipeline per shader ...
        assert [g for g, _, _ in groups] == sorted(bound)  # the same bind-group slots are populated
        for g, fld, idx in groups:
            key = (fld, idx)
            assert field_to_identity.setdefault(key, bound[g]) == bound[g], (label, g, key)
            assert identity_to_field.setdefault(bound[g], key) == key, (label, g, key)
        dispatch_fields.add(dispatch)
    assert len(set(pipeline_of.values())) == len(pipeline_of)  # ... and distinct shaders use distinct pipelines
    assert len(dispatch_fields) == 1  # every pass dispatches self.work_group_size = ceil(x*y / 256) workgroups
    return field_to_identity


def test_dispatch_order_equals_the_reference_binarys_command_stream(golden):
    """LBM::iterate of the reference's binary, run against a mock wgpu context, emits per step: one encoder with the
    passes Precollision-corner, Precollision-cardinal, Collision-corner, Collision-cardinal, one with Stream_e_w,
    Stream_n_s, Stream_nw_se, Stream_ne_sw, then (per call) the summary and colour-map passes.  The oracle-side
    driver must issue the same passes in the same order with a consistent one-to-one correspondence between bind
    groups: the population pairs alternate with compute_step % 2 (streams: source = step % 2, destination = the
    other), density / size / dimensions / barrier / output groups do not, and the collision group that carries the
    rest population is a single, parity-independent bind group (lbm.rs:775-778)."""
    import json
    m0 = check_command_stream(json.loads(golden["iterate_trace/steps3_from0"].item()), 3, 0, 0, 2)
    m7 = check_command_stream(json.loads(golden["iterate_trace/steps2_from7"].item()), 2, 7, 0, 2)
    assert m0 == m7  # the correspondence does not depend on where compute_step starts
    rest_groups = [k for k, v in m0.items() if ("rest",) in v]
    assert len(rest_groups) == 1 and rest_groups[0][1] is None  # one bind group, not indexed by parity
    pair_groups = [k for k, v in m0.items() if any(i[0] == "population" for i in v)]
    assert len(pair_groups) == 8 and all(idx in (0, 1) for _, idx in pair_groups)
    for (fld, idx), ident in m0.items():
        if idx is not None:  # element i of a Vec<BindGroup> holds the pair's arrays of data_buffers[i]
            assert {i[1] for i in ident} == {idx}
    pipelines = set()
    for stat in range(5):
        for cmap in range(3):
            stream = json.loads(golden[f"iterate_trace/frame_only/{stat}/{cmap}"].item())
            check_command_stream(stream, 0, 0, stat, cmap)
            pipelines.add(tuple(r[2] for r in stream if r[0] == "pass"))
    assert len(pipelines) == 15 and len({p[0] for p in pipelines}) == 5 and len({p[1] for p in pipelines}) == 3


def test_paint_command_stream_of_the_reference_binary(golden):
    """draw_shape run to completion against the mock wgpu context: two uploads (the pairs, then the count word), an
    empty submit, and ONE compute pass that binds (draw group, barrier group) and dispatches exactly one workgroup
    per painted point (barrier_draw.wgsl has @workgroup_size(1)) — what the oracle-side driver's draw_points does;
    and the barrier group it binds is the very group the four stream passes bind at slot 3."""
    import json
    stream = json.loads(golden["draw_trace/stream"].item())
    n = int(golden["draw_trace/npoints"])
    kinds = [r[0] for r in stream]
    assert kinds == ["write", "write", "submit", "encoder", "pass", "submit"]
    pairs = np.frombuffer(bytes.fromhex(stream[0][2]), dtype=np.uint32)
    count = np.frombuffer(bytes.fromhex(stream[1][2]), dtype=np.uint32)
    assert len(pairs) == 2 * n and count.tolist() == [2 * n - 1] and stream[0][1] != stream[1][1]
    _, label, pipeline, groups, dispatch = stream[4]
    assert dispatch == n and [g[0] for g in groups] == [0, 1] and all(g[2] is None for g in groups)
    step = json.loads(golden["iterate_trace/steps3_from0"].item())
    stream_passes = [r for r in step if r[0] == "pass" and r[1] and r[1].startswith("Stream_")]
    assert {tuple(r[3][3][1:]) for r in stream_passes} == {tuple(groups[1][1:])}  # the same barrier bind group
    assert pipeline not in {r[2] for r in step if r[0] == "pass"}  # its own pipeline (barrier_draw.wgsl)
    # the oracle-side driver: one dispatch of len(pairs) workgroups, group 0 = (count, updates), group 1 = barrier
    from oracle.wgsl_interp import WgslLBM

    class Recorder(WgslLBM):
        def _load_shaders(self, root):
            outer = self

            class Draw:
                def dispatch(self, workgroups, bindings):
                    outer.seen = (workgroups, {g for g, _ in bindings}, bindings[(1, 0)] is outer.barrier)
            self.sh = {"draw": Draw()}

    sim = Recorder(1.0, 64, 32)
    sim.draw_points(pairs.reshape(-1, 2))
    assert sim.seen == (n, {0, 1}, True)


@needs_reference
def test_fixture_comes_from_the_binary_in_the_reference_tree_and_reruns_live(golden):
    from tests.golden.make_wasm_golden import Reference
    ref = Reference()  # checks the fingerprints of the four functions
    assert ref.sha256 == golden["wasm_sha256"].item().decode()
    for ux, want in list(zip(golden["inflows"], golden["set_equil"]))[:3]:
        assert (bits(ref.set_equil(ux)) == bits(want)).all()
    assert (bits(ref.single_cell(2, 9, 7)) == bits(golden["single_cell/9x7/2"])).all()
    for i, p1, p2, xd, yd in list(cases(golden))[:4]:
        assert np.array_equal(np.array(ref.line_new(p1, p2, xd, yd), np.int32).reshape(-1, 3), golden[f"line/{i}/new"])
    i, p1, p2, xd, yd = list(cases(golden))[-1]
    assert ref.line_new(p1, p2, xd, yd) is None and len(golden[f"line/{i}/new"]) == 0


def test_wasm_interpreter_arithmetic():
    """the interpreter itself, on a hand-assembled module: i32 wrap-around, signed division and comparison,
    shifts, a loop with br_if, memory round trip, f32 rounding"""
    from oracle.wasm_mini import Instance, Module, Trap

    def leb(v):
        out = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            if (v == 0 and not b & 0x40) or (v == -1 and b & 0x40):
                return bytes(out + bytes([b]))
            out.append(b | 0x80)

    def func(locals_, body):
        code = bytes([len(locals_)]) + b"".join(bytes([n, t]) for n, t in locals_) + body + b"\x0b"
        return leb(len(code)) + code

    I, F = 0x7F, 0x7D
    # f0(a, b) = a / b (signed) + (a < b signed); f1(n) = sum_{k<n} k*k via a loop; f2(x, y) = x * y + 0.1 (f32), through memory
    f0 = func([], b"\x20\x00\x20\x01\x6d\x20\x00\x20\x01\x48\x6a")
    f1 = func([(2, I)], b"\x03\x40" + b"\x20\x01\x20\x02\x20\x02\x6c\x6a\x21\x01" + b"\x20\x02\x41\x01\x6a\x22\x02\x20\x00\x49\x0d\x00" + b"\x0b\x20\x01")
    f2 = func([], b"\x41\x10\x20\x00\x20\x01\x94\x38\x02\x00\x41\x10\x2a\x02\x00\x43" + np.float32(0.1).tobytes() + b"\x92")
    types = b"\x03" + b"\x60\x02\x7f\x7f\x01\x7f" + b"\x60\x01\x7f\x01\x7f" + b"\x60\x02\x7d\x7d\x01\x7d"
    sec = lambda sid, body: bytes([sid]) + leb(len(body)) + body  # noqa: E731
    mod = (b"\x00asm\x01\x00\x00\x00" + sec(1, types) + sec(3, b"\x03\x00\x01\x02") + sec(5, b"\x01\x00\x01") +
           sec(10, b"\x03" + f0 + f1 + f2))
    inst = Instance(Module(mod))
    assert inst.call(0, (-7) & 0xFFFFFFFF, 2) == ((-3) + 1) & 0xFFFFFFFF  # trunc toward zero, -7 < 2 signed
    assert inst.call(0, 7, (-2) & 0xFFFFFFFF) == (-3) & 0xFFFFFFFF
    with pytest.raises(Trap):
        inst.call(0, 1, 0)
    assert inst.call(1, 10) == sum(k * k for k in range(10))
    x, y = np.float32(1.1), np.float32(3.3)
    assert np.float32(inst.call(2, float(x), float(y))).view(np.uint32) == (x * y + np.float32(0.1)).view(np.uint32)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_initial_state_equals_the_reference_binarys_set_equil(golden):
    from lbm_b200 import LBM
    for ux, want in zip(golden["inflows"], golden["set_equil"]):
        lbm = LBM(1.0, 70, 9, inflow_ux=float(ux))
        for b in (0, 1):
            for k in range(9):
                assert (bits(lbm.read_population(k, b)) == bits(want[k])).all(), (ux, b, k)
        lbm.custom_speed(float(ux))
        for k in range(9):
            assert (bits(lbm.read_population(k, 0)) == bits(want[k])).all(), ("custom_speed", ux, k)
        lbm.close()


@pytest.mark.gpu
def test_cuda_single_cell_equals_what_the_reference_binary_uploads(golden):
    from lbm_b200 import LBM
    for x, y, index, want in single_cell_cases(golden):
        lbm = LBM(1.0, x, y)
        lbm.iterate(3)
        lbm.single_cell(index)
        for b in (0, 1):
            for k in range(9):
                assert (bits(lbm.read_population(k, b)) == bits(want[b, k])).all(), (x, y, index, b, k)
        assert lbm.get_compute_num() == 0
        lbm.close()


@pytest.mark.gpu
def test_cuda_draw_points_accepts_what_the_reference_binary_uploads(golden):
    """the drop-in check for row a-7: the very arrays the reference's draw_shape uploads, handed to blbm_draw_points"""
    from lbm_b200 import LBM
    n = 0
    for p1, p2, xd, yd, erase, pairs, count, pts in draw_cases(golden):
        lbm = LBM(1.0, xd, yd)
        if erase:
            lbm.draw_points(np.stack([pairs[:, 0], np.ones_like(pairs[:, 0])], 1))  # something to erase
        expect = lbm.read_barrier().copy()
        lbm.draw_points(pairs.astype(np.uint32))
        expect[pts[:, 1], pts[:, 0]] = 0 if erase else 1
        assert np.array_equal(lbm.read_barrier(), expect), (p1, p2, erase)
        lbm.close()
        n += 1
    assert n >= 16


@pytest.mark.gpu
def test_cuda_draw_line_paints_the_reference_binarys_cells(golden):
    """blbm_draw_line / blbm_erase_line (draw_shape(&Line::new(..)) / draw_shape(&Line::new_erased(..))) on the device
    mask against the cells the reference's binary produced"""
    from lbm_b200 import LBM
    done = 0
    for i, p1, p2, xd, yd in cases(golden):
        want = golden[f"line/{i}/new"]
        if xd != 64 or len(want) == 0 or i % 2:
            continue
        lbm = LBM(1.0, xd, yd)
        base = lbm.read_barrier().copy()
        lbm.draw_line(p1, p2)
        expect = base.copy()
        expect[want[:, 1], want[:, 0]] = 1
        assert np.array_equal(lbm.read_barrier(), expect), f"draw_line {p1}->{p2}"
        key = f"line/{i}/erased"
        if key in golden.files:
            er = golden[key]
            lbm.erase_line(p1, p2)
            expect[er[:, 1], er[:, 0]] = 0
            assert np.array_equal(lbm.read_barrier(), expect), f"erase_line {p1}->{p2}"
        lbm.close()
        done += 1
    assert done >= 10


# ---- the barrier presets, pinned to the `match *BARRIER_PRESET` arms of the binary's event loop ----------------------
PRESETS = os.path.join(HERE, "golden", "wasm_presets.npz")
PRESET_IDS = {"curl": 0, "chaos": 1, "welcome": 2}


@pytest.fixture(scope="module")
def presets():
    return np.load(PRESETS)


def _preset_draws(presets, name, x, y):
    out, k = [], 0
    while f"{name}/{x}x{y}/draw{k}" in presets:
        out.append(presets[f"{name}/{x}x{y}/draw{k}"])
        k += 1
    return out


def test_preset_end_points_equal_the_reference_binarys_line_new_calls(presets, golden):
    """LBM::curl_barrier / chaos_barrier / welcome_barrier (lbm.rs:1367-1480) are inlined into the event-loop closure of
    the shipped binary; the fixture holds what that code passed to Line::new when the arm itself was executed.  The
    product's end-point arithmetic (blbm_preset_lines, pure host code) must reproduce every argument in order."""
    from lbm_b200.lbm import preset_lines
    assert bytes(presets["wasm_sha256"]) == bytes(golden["wasm_sha256"])
    n = 0
    for x, y in presets["sizes"]:
        x, y = int(x), int(y)
        for name, pid in PRESET_IDS.items():
            want = presets[f"{name}/{x}x{y}/lines"]
            assert (want[:, 4] == x).all() and (want[:, 5] == y).all()
            np.testing.assert_array_equal(preset_lines(pid, x, y), want[:, :4], err_msg=f"{name} on {x}x{y}")
            n += len(want)
    assert n == 5 * (1 + 4 + 28)


def test_preset_masks_equal_what_the_reference_binary_draws(presets):
    """the point sets the binary hands to draw_shape, preset by preset (chaos: four separate draws; welcome: one
    Blob), against the oracle-side restatement and against the product's rasteriser applied to its own end points"""
    from lbm_b200.lbm import preset_lines, rasterize_line
    for x, y in presets["sizes"]:
        x, y = int(x), int(y)
        for name, pid in PRESET_IDS.items():
            draws = _preset_draws(presets, name, x, y)
            assert all((d[:, 2] == 1).all() for d in draws)
            want = {(int(a), int(b)) for d in draws for a, b, _ in d}
            got = {(p[0], p[1]) for p in getattr(barrier_shapes, name + "_barrier")(x, y)}
            assert got == want, f"oracle {name} on {x}x{y}"
            mine = set()
            per_line = []
            for x1, y1, x2, y2 in preset_lines(pid, x, y):
                pts = rasterize_line((x1, y1), (x2, y2), x, y)
                per_line.append({(int(a), int(b)) for a, b in pts})
                mine |= per_line[-1]
            assert mine == want, f"product {name} on {x}x{y}"
            if name == "chaos":  # drawn line by line
                for d, pl in zip(draws, per_line):
                    assert {(int(a), int(b)) for a, b, _ in d} == pl


@needs_reference
def test_preset_fixture_reruns_live(presets):
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_wasm_golden as gen
    ref = gen.Reference()
    assert ref.sha256.encode() == bytes(presets["wasm_sha256"])
    for name in ("curl", "chaos"):
        lines, draws = gen.preset_trace(ref, name, 300, 170)
        np.testing.assert_array_equal(np.array(lines), presets[f"{name}/300x170/lines"])
        for k, d in enumerate(draws):
            np.testing.assert_array_equal(np.array(d, np.int32).reshape(-1, 3), presets[f"{name}/300x170/draw{k}"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["curl", "chaos", "welcome"])
def test_cuda_presets_paint_what_the_reference_binary_draws(presets, name):
    from lbm_b200 import LBM
    for x, y in ((300, 170), (512, 256), (1000, 500)):
        lbm = LBM(1.25, x, y)
        getattr(lbm, name + "_barrier")()
        want = np.zeros((y, x), np.uint32)
        want[0] = want[-1] = 1
        for d in _preset_draws(presets, name, x, y):
            want[d[:, 1], d[:, 0]] = d[:, 2]
        np.testing.assert_array_equal(lbm.read_barrier(), want, err_msg=f"{name} on {x}x{y}")
        lbm.iterate(3)  # and the lattice steps with it
        lbm.close()


# ---- the CONTENTS of the bind groups, from LBM::new as the binary itself runs it ---------------------------------------
BINDGROUPS = os.path.join(HERE, "golden", "wasm_bindgroups.json")


def ordered_driver_bindings(n, compute_step, stat, cmap):
    """like recorded_driver, but per (group, binding): [(shader name, {(group, binding): identity})]"""
    from oracle.wgsl_interp import REST, WgslLBM

    class Recorder(WgslLBM):
        def _load_shaders(self, root):
            self.sh = {}

        def identity(self, b):
            for buf in (0, 1):
                for k in range(9):
                    if b is self.data[buf][k]:
                        return ("rest",) if k == REST else ("population", buf, k)
            for nm in ("ux", "uy", "rho", "output", "barrier", "colors"):
                if b is getattr(self, nm):
                    return (nm,)
            return ("dims",) if isinstance(b, dict) else ("size",) if isinstance(b, np.uint32) else ("omega",)

        def _run(self, name, bindings):
            self.log.append((name, {gb: self.identity(b) for gb, b in bindings.items()}))

        def color_map(self, name=None):
            self._run(name or self.color_map_name, {(0, 0): self.colors, (1, 0): self.output, (2, 0): self.barrier,
                                                    (3, 0): np.uint32(self.n)})

    sim = Recorder(1.0, 8, 4)
    sim.log = []
    sim.compute_step = compute_step
    sim.set_summary(Recorder.STATS[stat])
    sim.iterate(n)
    sim.color_map(Recorder.CMAPS[cmap])
    return sim.log


def test_bind_group_contents_equal_what_the_reference_binary_builds(golden):
    """Which buffer sits at which binding of which bind group — `LBM::new`, lbm.rs:726-1049 — is no longer restated
    by hand.  tests/golden/wasm_bindgroups.json is what the binary's own LBM::new did when it was executed inside the
    start-up closure (its wgpu object creations served by the host, every object tagged): the buffers in creation order
    with their contents, the entries of all bind groups, and where each buffer / bind group ends up in the assembled
    LBM value.  Chained with the command stream of the binary's own LBM::iterate (which field it binds to which slot
    of which pass), that gives, from the binary alone, the buffer behind every (group, binding) of every pass — which
    must be what the oracle-side driver (the dispatch code all WGSL golden vectors were produced with) binds."""
    import json
    bg = json.load(open(BINDGROUPS))
    assert bg["wasm_sha256"].encode() == bytes(golden["wasm_sha256"])
    x, y = bg["x"], bg["y"]
    bufs = bg["buffers"]
    # -- what each buffer is, from its own creation: position in data_buffers, or contents
    ident = {}
    init = Oracle(1.25, 4, 4)  # any handle: set_equil lives in the library
    from oracle.lbm_oracle import set_equil
    eq = set_equil(0.1, 0.0, 1.0)
    init.close()
    for b in range(2):
        for k in range(9):
            n = bg["data_buffers"][b][k]
            # LBM::new fills both sets from set_equil(0.1, 0, 1): x*y copies of the k-th value (lbm.rs:739-744)
            assert bufs[n]["bytes"] == 4 * x * y and bufs[n]["uniform_contents"]
            assert bufs[n]["first_word"] == int(np.float32(eq[k]).view(np.uint32)), (b, k)
            ident[n] = ("rest",) if k == 4 and b == 0 else ("population", b, k)
    assert bg["data_buffers"] == [list(range(9)), list(range(9, 18))]  # created set by set, k = 0..8
    mask = np.zeros((y, x), np.uint32)
    mask[0] = mask[-1] = 1
    import hashlib
    barrier = [b["index"] for b in bufs if b.get("sha256") == hashlib.sha256(mask.tobytes()).hexdigest()]
    omega = [b["index"] for b in bufs if b.get("bytes") == 4 and b.get("usage") == 72 and b["index"] < 28]
    size = [b["index"] for b in bufs if b.get("bytes") == 4 and b.get("first_word") == x * y]
    dims = [b["index"] for b in bufs if b.get("sha256") == hashlib.sha256(np.array([x, y, x * y], np.uint32).tobytes()).hexdigest()]
    assert len(barrier) == 1 and len(omega) == 1 and len(size) == 1 and len(dims) == 2  # compute + vertex copies
    ident.update({barrier[0]: ("barrier",), omega[0]: ("omega",), size[0]: ("size",), dims[0]: ("dims",), dims[1]: ("dims",)})
    groups = {g["index"]: g for g in bg["bind_groups"]}
    # the zero-initialised storage buffers are told apart by the bind group LBM::new puts them in (lbm.rs:780-786:
    # density_bg = three of them, output_bg = one, color_bg = one) and named as the shaders declare the bindings
    fields = {int(k): v for k, v in bg["lbm_fields"].items()}
    vecs = {int(k): v for k, v in bg["lbm_bind_group_vecs"].items()}
    # -- chain with the command stream of LBM::iterate: per pass, slot -> LBM field -> bind group -> entries
    stream = (json.loads(bytes(golden["iterate_trace/steps3_from0"]).decode()) +
              json.loads(bytes(golden["iterate_trace/steps2_from7"]).decode()))
    log = ordered_driver_bindings(3, 0, 0, 2) + ordered_driver_bindings(2, 7, 0, 2)
    passes = [r for r in stream if r[0] == "pass"]
    assert len(passes) == len(log) == 8 * 3 + 2 + 8 * 2 + 2
    names = {}  # zero buffers: named on first sight by what the driver calls that (group, binding) ...
    seen_groups = set()
    for (_, label, _, slots, _), (shader, bound) in zip(passes, log):
        got = {}
        for slot, fld, idx in slots:
            if idx is None:
                kind, gi = fields[fld]
                assert kind == "bind_group", (label, slot, fld)
            else:
                gi = vecs[fld][idx]
            seen_groups.add(gi)
            for binding, n in groups[gi]["entries"]:
                who = ident.get(n)
                if who is None:
                    who = names.setdefault(n, bound[(slot, binding)])  # ... and must keep that name in every pass
                got[(slot, binding)] = who
        assert got == bound, (label or shader, got, bound)
    assert sorted(names.values()) == [("colors",), ("output",), ("rho",), ("ux",), ("uy",)] and len(names) == 5
    # the rest population is bound from buffer set 0 only (lbm.rs:775-778): data_buffers[1][4] is in no bind group
    dead = bg["data_buffers"][1][4]
    assert all(n != dead for g in groups.values() for _, n in g["entries"])
    assert [e for e in groups[8]["entries"]] == [[0, size[0]], [1, omega[0]], [2, bg["data_buffers"][0][4]]]
    # every bind group LBM::iterate uses was reached (the draw group and the vertex dimensions belong to paint / render)
    assert seen_groups == set(range(16)) - {14}


@needs_reference
def test_bind_group_fixture_reruns_live(golden):
    import json
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_wasm_golden as gen
    ref = gen.Reference()
    live = gen.bindgroup_trace(ref)
    stored = json.load(open(BINDGROUPS))
    for key in ("buffers", "bind_groups", "lbm_fields", "lbm_bind_group_vecs", "data_buffers", "created"):
        assert live[key] == stored[key], key
