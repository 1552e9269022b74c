"""TEST-DEVELOPMENT TOOL, not a test and not a parity statement: runs the GPU-marked tests that were written while no
GPU was at hand (tests/test_wasm_pin.py, tests/test_wgsl_pin.py: executed-WGSL digests, wasm pins) against a MOCK of
lbm_b200.LBM built on the C oracle, to catch Python-level mistakes in the test code itself (names, shapes, keys,
fixture entries) before the tests meet a B200.  It patches lbm_b200.LBM inside this process only.

    python tests/mock_gpu_dryrun.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.lbm_oracle import Oracle
from oracle import barrier_shapes
import lbm_b200, lbm_b200.lbm as host

class MockLBM:
    def __init__(self, omega, x, y, inflow_ux=0.1, device=0, rows=None, kernel=None, lazy_barriers=None):
        self.x, self.y = x, y
        self.o = Oracle(omega, x, y, inflow_ux=inflow_ux)
    def __getattr__(self, name):
        return getattr(self.o, name)
    def read_population(self, k, buffer=-1):
        b = self.o.get_compute_num() % 2 if buffer < 0 else buffer
        return np.array(self.o.population(b, k)).reshape(self.y, self.x)
    def read_moments(self):
        return tuple(np.array(a).reshape(self.y, self.x) for a in self.o.moments())
    def read_output(self):
        return np.array(self.o.output()).reshape(self.y, self.x)
    def read_barrier(self):
        return np.array(self.o.barrier()).reshape(self.y, self.x)
    def _line(self, p1, p2, erase):
        pts = barrier_shapes.line_points(p1, p2, self.x, self.y, erase=erase)
        a = np.array([[px + py * self.x, 0 if erase else 1] for px, py, *_ in pts], np.uint32)
        self.o.draw_points(a)
    def color_map(self, cmap): self._rgb = self.o.color_map(cmap)
    def read_colors(self): return self._rgb
    def draw_line(self, p1, p2): self._line(p1, p2, False)
    def erase_line(self, p1, p2): self._line(p1, p2, True)
    def close(self): self.o.close()

class MockGroup(MockLBM):
    def __init__(self, omega, x, y, devices, inflow_ux=0.1, kernel=None, lazy_barriers=None):
        super().__init__(omega, x, y, inflow_ux=inflow_ux)

lbm_b200.LBM = MockLBM
host.LBM = MockLBM
host.SlabGroup = MockGroup

import tests.test_wasm_pin as tw
import tests.test_wgsl_pin as tg
g = np.load(tw.GOLDEN)
tw.test_cuda_initial_state_equals_the_reference_binarys_set_equil(g); print("ok set_equil")
tw.test_cuda_single_cell_equals_what_the_reference_binary_uploads(g); print("ok single_cell")
tw.test_cuda_draw_line_paints_the_reference_binarys_cells(g); print("ok draw_line")
tw.test_cuda_draw_points_accepts_what_the_reference_binary_uploads(g); print("ok draw_points wire format")
c1 = np.load(tg.CONFIG1); wide = np.load(tg.WIDE); gold = np.load(tg.GOLDEN)
skip = {"box_4096x4096", "box_4096x4096_1000steps"}
for name in sorted(tg.WIDE_CASES):
    if name in skip: continue
    tg.test_cuda_reproduces_the_executed_reference_shaders_on_large_lattices(wide, name, 2); print("ok wide", name)
for name, n in (("random_300x170", 4), ("random_1001x37", 2)):
    tg.test_cuda_slab_group_reproduces_the_executed_reference_shaders(wide, name, n, 1); print("ok slabs", name)
tg.test_cuda_reproduces_config1_10k_steps_of_the_executed_reference_shaders(c1, 0.1, 2); print("ok config1")
