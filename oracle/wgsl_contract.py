"""TEST INFRASTRUCTURE — the contraction twin, derived from the shader text instead of by hand.

WGSL allows a backend to fuse a floating-point multiply into the add or subtract that consumes it (naga emits no
NoContraction).  The default oracle semantics are "no contraction"; the twin builds (oracle -DLBM_CONTRACT, CUDA
-DBLBM_CONTRACT) fuse a fixed list of pairs.  This module derives that list mechanically from the reference's WGSL
text and executes the shaders under it, so that the twin is pinned to the shader source exactly like the default
mode is:

  * `analyse(shader)`: value-numbers the @compute main (common subexpressions share a number, `let` names are
    transparent, storage loads are versioned by the stores in between), counts the uses of every value, and applies
    the rule an LLVM-style backend applies after CSE: in `a + b` / `a - b`, a float multiply operand with exactly ONE
    use is fused — the left operand first (`fadd (fmul x y) z -> fma x y z`, then `fadd x (fmul y z) -> fma y z x`;
    `fsub (fmul x y) z -> fma x y (-z)`, `fsub x (fmul y z) -> fma (-y) z x`).
  * `ContractLanes`: the SIMT executor of wgsl_simt.py with those additions evaluated as a correctly rounded fp32 fma
    (exact product in f64, sum rounded to odd in f64, then to f32 — f64 carries 53 >= 2*24 + 2 bits).

Only tests/ and golden generators may import this; the product never does.
"""
import os

import numpy as np

from .wgsl_simt import VecShader, WgslLBMVec, _Lanes

f32, f64 = np.float32, np.float64


# ---------------------------------------------------------------------------------------------------------------
# analysis: which additions / subtractions of a shader fuse which multiply
# ---------------------------------------------------------------------------------------------------------------
class _Numbering:
    def __init__(self, shader):
        self.sh = shader
        self.nodes = {}    # key -> value number
        self.info = []     # value number -> (key, is_float)
        self.uses = []     # value number -> number of users (distinct parent values, stores, conditions)
        self.version = {}  # array name -> number of stores so far (loads across a store are different values)
        self.adds = []     # (ast node, vn of lhs, vn of rhs) for every float + / -
        self.float_arrays = {n for n, g in shader.globals.items() if g["type"] == ("array", ("f32", None))}

    def vn(self, key, is_float, operands=()):
        if key not in self.nodes:
            self.nodes[key] = len(self.info)
            self.info.append((key, is_float))
            self.uses.append(0)
            for o in operands:  # a value is used once by every distinct value built from it
                self.uses[o] += 1
        return self.nodes[key]

    def use(self, v):
        self.uses[v] += 1

    def is_float(self, v):
        return self.info[v][1]

    def expr(self, e, env):
        kind = e[0]
        if kind == "lit":
            return self.vn(("lit", e[1], e[2]), e[2] in ("abstract-float", "f32"))
        if kind == "var":
            if e[1] in env:
                return env[e[1]]
            g = self.sh.globals.get(e[1])
            return self.vn(("global", e[1]), bool(g) and g["type"] == ("f32", None))
        if kind == "member":
            b = self.expr(e[1], env)
            return self.vn(("member", b, e[2]), False, (b,))
        if kind == "index":
            assert e[1][0] == "var"
            i = self.expr(e[2], env)
            name = e[1][1]
            return self.vn(("load", name, i, self.version.get(name, 0)), name in self.float_arrays, (i,))
        if kind in ("neg", "not"):
            a = self.expr(e[1], env)
            return self.vn((kind, a), self.is_float(a), (a,))
        if kind == "call":
            args = tuple(self.expr(a, env) for a in e[2])
            if e[1] in ("f32",):
                fl = True
            elif e[1] in ("u32", "i32"):
                fl = False
            else:
                fl = any(self.is_float(a) for a in args) or e[1].startswith("vec")
            return self.vn(("call", e[1]) + args, fl, args)
        if kind == "bin":
            op = e[1]
            a, b = self.expr(e[2], env), self.expr(e[3], env)
            ka, kb = self.info[a][0], self.info[b][0]
            if ka[0] == "lit" and kb[0] == "lit" and ka[2].startswith("abstract") and kb[2].startswith("abstract"):
                # a constant expression of abstract literals (1.0/36.0) is folded by the front end: a literal
                val = {"+": ka[1] + kb[1], "-": ka[1] - kb[1], "*": ka[1] * kb[1], "/": ka[1] / kb[1]}[op]
                return self.vn(("lit", val, "abstract-float"), True)
            fl = (self.is_float(a) or self.is_float(b)) and op in "+-*/"
            v = self.vn(("bin", op, a, b), fl, (a, b))
            if fl and op in "+-":
                self.adds.append((e, a, b))
            return v
        raise NotImplementedError(kind)

    def block(self, stmts, env):
        env = dict(env)
        for s in stmts:
            kind = s[0]
            if kind in ("let", "var"):
                env[s[1]] = self.expr(s[3], env) if s[3] is not None else self.vn(("undef", id(s)), False)
            elif kind == "assign":
                lhs, rhs = s[2], s[3]
                if s[1] != "=":  # a += b is a = a + b: one more float addition
                    rhs = ("bin", s[1][0], lhs, rhs)
                    self.compound[id(s)] = rhs
                v = self.expr(rhs, env)
                self.use(v)
                if lhs[0] == "index":
                    self.use(self.expr(lhs[2], env))
                    self.version[lhs[1][1]] = self.version.get(lhs[1][1], 0) + 1
                elif lhs[0] == "var":
                    env[lhs[1]] = v
                else:
                    raise NotImplementedError(lhs[0])
            elif kind == "if":
                self.use(self.expr(s[1], env))
                self.block(s[2], env)
                if s[3]:
                    self.block(s[3], env)
            elif kind == "switch":
                self.use(self.expr(s[1], env))
                for _, body in s[2]:
                    self.block(body, env)
                if s[3]:
                    self.block(s[3], env)
            elif kind == "block":
                self.block(s[1], env)
            elif kind == "return":
                if s[1] is not None:
                    self.use(self.expr(s[1], env))
            elif kind == "expr":
                self.use(self.expr(s[1], env))
            else:
                raise NotImplementedError(kind)


def analyse(shader):
    """{id(ast node of a float + or -): 'lhs' | 'rhs'} — which operand (a single-use multiply) is fused into it; plus
    the rewritten right-hand sides of compound assignments (their additions exist only after `a += b` -> `a = a + b`)
    and a readable list of the fusions."""
    n = _Numbering(shader)
    n.compound = {}
    fn = shader.funcs["main"]
    env = {p[0]: n.vn(("param", p[0]), False) for p in fn["params"]}
    n.block(fn["body"], env)

    def single_use_mul(v):
        key, fl = n.info[v]
        return fl and key[0] == "bin" and key[1] == "*" and n.uses[v] == 1

    fused, listing = {}, []
    for ast, a, b in n.adds:
        if id(ast) in fused:
            continue
        side = "lhs" if single_use_mul(a) else "rhs" if single_use_mul(b) else None
        if side:
            fused[id(ast)] = side
            listing.append((ast[1], side))
    return fused, n.compound, listing


# ---------------------------------------------------------------------------------------------------------------
# execution under the derived fusions
# ---------------------------------------------------------------------------------------------------------------
def fma32(a, b, c):
    """correctly rounded fp32 fma on numpy values: the product of two f32 is exact in f64; the sum is rounded to odd in
    f64 (TwoSum gives the rounding error), which makes the final rounding to f32 the rounding of the exact result"""
    a, b, c = (np.asarray(v, f32).astype(f64) for v in (a, b, c))
    p = a * b
    s = p + c
    bb = s - p
    err = (p - (s - bb)) + (c - bb)
    inexact = (err != 0) & np.isfinite(s)
    even = (s.view(np.int64) & 1) == 0
    toward = np.where((err > 0), np.inf, -np.inf)
    s = np.where(inexact & even, np.nextafter(s, toward), s)
    return s.astype(f32)


class ContractShader(VecShader):
    def __init__(self, path):
        super().__init__(path)
        self.fused, self.compound, self.listing = analyse(self)


class ContractLanes(_Lanes):
    """_Lanes with the fused additions evaluated as fma.  A multiply that reaches its addition through a `let` name
    hands over its operand VALUES (recorded when the let was executed), so the fma sees exactly what the multiply saw."""

    def _stmt(self, s, fr):
        if s[0] == "let" and s[3] is not None and s[3][0] == "bin" and s[3][1] == "*":
            x, y = self._eval(s[3][2], fr), self._eval(s[3][3], fr)
            fr.scopes[-1][s[1]] = self._binop("*", x, y)
            self.mulops[s[1]] = (x, y)
            return
        if s[0] == "assign" and id(s) in self.sh.compound:
            return super()._stmt(("assign", "=", s[2], self.sh.compound[id(s)]), fr)
        return super()._stmt(s, fr)

    def run(self):
        self.mulops = {}
        return super().run()

    def _mul_operands(self, e, fr):
        if e[0] == "var" and e[1] in self.mulops:
            return self.mulops[e[1]]
        assert e[0] == "bin" and e[1] == "*", e
        return self._eval(e[2], fr), self._eval(e[3], fr)

    def _eval(self, e, fr):
        if e[0] == "bin" and id(e) in self.sh.fused:
            side, op = self.sh.fused[id(e)], e[1]
            if side == "lhs":
                x, y = self._mul_operands(e[2], fr)
                z = self._eval(e[3], fr)
                return self._fma(x, y, z if op == "+" else self._neg(z))
            x, y = self._mul_operands(e[3], fr)
            z = self._eval(e[2], fr)
            return self._fma(x if op == "+" else self._neg(x), y, z)
        return super()._eval(e, fr)

    @staticmethod
    def _neg(v):
        from .wgsl_simt import Vec3
        return Vec3([-c for c in v.c]) if isinstance(v, Vec3) else -np.asarray(v, f32)

    @staticmethod
    def _fma(x, y, z):
        from .wgsl_simt import Vec3
        if any(isinstance(v, Vec3) for v in (x, y, z)):
            comp = [v.c if isinstance(v, Vec3) else [v] * 3 for v in (x, y, z)]
            return Vec3([fma32(a, b, c) for a, b, c in zip(*comp)])
        return fma32(x, y, z)


def _dispatch_contract(shader, workgroups, bindings):
    """VecShader.dispatch with ContractLanes (single chunk)"""
    from .wgsl_simt import _Dispatch
    n = workgroups * shader.workgroup_size
    arrays, env = {}, {}
    for name, g in shader.globals.items():
        b = bindings[(g["group"], g["binding"])]
        if isinstance(b, np.ndarray):
            arrays[name] = b
            env[name] = ("array", name)
        else:
            env[name] = b
    ContractLanes(shader, _Dispatch(arrays, env, shader.path), 0, n).run()


class WgslLBMContract(WgslLBMVec):
    """the reference's host-side dispatch order with every shader executed under the derived contraction"""

    def __init__(self, omega, x, y, inflow_ux=0.1, root=None):
        super().__init__(omega, x, y, inflow_ux=inflow_ux, root=root, threads=1)
        self.vsh = {k: ContractShader(v.path) for k, v in self.vsh.items()}

    def color_map(self, name=None):
        name = name or self.color_map_name
        _dispatch_contract(self.vsh[name], self.work_groups, {(0, 0): self.colors, (1, 0): self.output,
                                                              (2, 0): self.barrier, (3, 0): np.uint32(self.n)})

    def _run(self, name, bindings):
        if name in self.vsh:
            _dispatch_contract(self.vsh[name], self.work_groups, bindings)
        else:
            self.sh[name].dispatch(self.work_groups, bindings)


def fusion_table(root=None):
    """{shader file: [(operator, fused side), ...]} for every shader of the reference, in source order"""
    from . import wgsl_interp
    root = root or wgsl_interp.SHADER_ROOT
    out = {}
    for sub in sorted(os.listdir(root)):
        d = os.path.join(root, sub)
        if not os.path.isdir(d):
            continue
        for f in sorted(os.listdir(d)):
            if f.endswith(".wgsl"):
                try:
                    out[f"{sub}/{f}"] = ContractShader(os.path.join(d, f)).listing
                except (SyntaxError, NotImplementedError, AssertionError) as e:  # barrier_erase.wgsl etc.
                    out[f"{sub}/{f}"] = f"not analysed: {e}"
    return out
