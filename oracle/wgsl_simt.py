"""TEST INFRASTRUCTURE — a SIMT executor for the reference's WGSL compute shaders: the same parser and the same
semantics as oracle/wgsl_interp.py, but all invocations of a dispatch are evaluated together on numpy arrays
under an execution mask, the way a GPU runs them.  That makes the reference's shader text executable at the
reference's own problem size: BASELINE.json configs[0] (512 x 256, 10,000 steps = 80,000 dispatches of 131,072
invocations) takes minutes instead of days, so the 10k-step parity target of the north star is pinned to the
reference's source and not to a restatement (tests/golden/make_wgsl_config1.py, tests/test_wgsl_pin.py).

Why evaluating a dispatch "all lanes at once" is the same as any sequential or parallel order: the executor
REFUSES a shader unless, within one dispatch, every store goes to the invocation's own element (index ==
global_invocation_id.x) and every load from an array that the dispatch also stores to — before or after the load
in program order — reads the invocation's own element (`_check_own`, `_Dispatch.stored / foreign`).  Then no
invocation can observe another one's stores and the result does not depend on scheduling; for the same reason
the invocations may be split into contiguous chunks that run on several host threads (`threads=`; numpy releases
the GIL inside array operations).  All step, summary and colour-map shaders of the reference have that property; the
scatter of barrier_draw.wgsl stays with the scalar interpreter.

Semantics (identical to wgsl_interp.py and to the oracle's normative block): every fp32
operation individually rounded (numpy float32 element-wise arithmetic; IEEE division and sqrt), u32 arithmetic
wraps, out-of-range reads return 0, out-of-range writes are dropped, abstract literal expressions are evaluated in
f64 / integers and take the concrete type of the operand they meet.

Only tests/ and the golden generators may import this; the product never does.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .wgsl_interp import NW, N, NE, W, REST, E, SW, S, SE, Parser, Shader, WgslLBM, _is_abstract, f32, i32, u32  # noqa: F401

_CONCRETE = {"u32": u32, "i32": i32, "f32": f32}


def _concretise(v, like):
    if isinstance(like, np.ndarray):
        return like.dtype.type(v)
    return type(like)(v)


def _unify(a, b):
    if _is_abstract(a) and not _is_abstract(b):
        a = _concretise(a, b)
    elif _is_abstract(b) and not _is_abstract(a):
        b = _concretise(b, a)
    elif _is_abstract(a) and _is_abstract(b):
        if isinstance(a, float) or isinstance(b, float):
            a, b = float(a), float(b)
    return a, b


def _is_int(v):
    if isinstance(v, np.ndarray):
        return v.dtype.kind in "iu"
    return isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_))


OOB_POLICY = "zero"  # "clamp": sensitivity study of the other legal out-of-range-read behaviour (see _Lanes._load)


class Vec3:
    """vec3<T>: three components, each a per-lane array or a uniform / abstract scalar"""

    def __init__(self, comps):
        self.c = list(comps)


class _Frame:
    """one function activation: lexical scopes, the lanes still running, the return value being assembled"""

    def __init__(self, scopes, active):
        self.scopes = scopes
        self.active = active
        self.ret = None


class _Dispatch:
    """what the chunks of one dispatch share: the bound arrays and the record of which arrays were stored to and
    which were read at another invocation's element (set operations are atomic under the GIL; every access first
    records itself and then looks at the other set, so of two conflicting accesses at least one sees the other)"""

    def __init__(self, arrays, env, path):
        self.arrays, self.env, self.path = arrays, env, path
        self.stored, self.foreign = set(), set()

    def conflict(self, name):
        return RuntimeError(f"SIMT executor: `{name}` is stored to and also read or written at another invocation's "
                            f"element in {os.path.basename(self.path)}: the result would depend on scheduling")


class VecShader:
    def __init__(self, path):
        self.path = path
        self.structs, self.globals, self.funcs = Parser(open(path).read()).module()
        if "main" not in self.funcs or "compute" not in self.funcs["main"]["attrs"]:
            raise SyntaxError("WGSL: no @compute fn main")
        self.workgroup_size = int(self.funcs["main"]["attrs"]["workgroup_size"][0])

    # bindings as for Shader.dispatch: {(group, binding): numpy array | numpy scalar | dict (uniform struct)}
    def dispatch(self, workgroups, bindings, threads=1, pool=None):
        n = workgroups * self.workgroup_size
        arrays, env = {}, {}
        for name, g in self.globals.items():
            key = (g["group"], g["binding"])
            if key not in bindings:
                raise KeyError(f"WGSL: nothing bound at @group({key[0]}) @binding({key[1]}) for `{name}`")
            b = bindings[key]
            if (g["type"][0] == "array") != isinstance(b, np.ndarray):
                raise TypeError(f"WGSL: binding for `{name}` has the wrong shape")
            if isinstance(b, np.ndarray):
                if b.dtype.names:
                    raise NotImplementedError("SIMT executor: arrays of structs")
                arrays[name] = b
                env[name] = ("array", name)
            else:
                env[name] = b
        d = _Dispatch(arrays, env, self.path)
        threads = max(1, min(int(threads), n // 65536 or 1))
        if threads == 1:
            _Lanes(self, d, 0, n).run()
            return
        step = -(-n // threads)
        chunks = [_Lanes(self, d, lo, min(lo + step, n)) for lo in range(0, n, step)]
        own = pool is None
        pool = pool or ThreadPoolExecutor(threads)
        try:
            for f in [pool.submit(c.run) for c in chunks]:
                f.result()
        finally:
            if own:
                pool.shutdown()


class _Lanes:
    """the invocations [lo, hi) of one dispatch, evaluated together"""

    def __init__(self, shader, dispatch, lo, hi):
        self.sh, self.d, self.lo, self.hi = shader, dispatch, lo, hi
        self.n = hi - lo
        self.gid = np.arange(lo, hi, dtype=u32)

    def run(self):
        fn = self.sh.funcs["main"]
        scope = {}
        for pname, _, pattrs in fn["params"]:
            if pattrs.get("builtin") == ["global_invocation_id"]:
                scope[pname] = {"x": self.gid, "y": u32(0), "z": u32(0)}
            else:
                raise NotImplementedError(f"WGSL: builtin {pattrs}")
        with np.errstate(all="ignore"):
            self._block(fn["body"], _Frame([self.d.env, scope], np.ones(self.n, dtype=bool)))

    # -- statements ------------------------------------------------------------------------------------------
    def _block(self, stmts, fr):
        fr.scopes.append({})
        for s in stmts:
            if not fr.active.any():
                break
            self._stmt(s, fr)
        fr.scopes.pop()

    def _stmt(self, s, fr):
        kind = s[0]
        if kind == "return":
            if s[1] is not None:
                v = self._eval(s[1], fr)
                if fr.ret is None:
                    fr.ret = v if not isinstance(v, np.ndarray) else v.copy()
                else:
                    fr.ret = np.where(fr.active, v, fr.ret)
            fr.active = np.zeros_like(fr.active)
        elif kind == "block":
            self._block(s[1], fr)
        elif kind == "if":
            cond = self._eval(s[1], fr)
            if not isinstance(cond, np.ndarray):
                cond = np.full(self.n, bool(cond))
            before = fr.active
            taken, not_taken = before & cond, before & ~cond
            after = np.zeros_like(before)
            if taken.any():
                fr.active = taken
                self._block(s[2], fr)
                after |= fr.active
            if s[3] is not None and not_taken.any():
                fr.active = not_taken
                self._block(s[3], fr)
                after |= fr.active
            elif s[3] is None:
                after |= not_taken
            fr.active = after
        elif kind == "switch":
            sel = self._eval(s[1], fr)
            before = fr.active
            matched = np.zeros_like(before)
            after = np.zeros_like(before)
            for vals, body in s[2]:
                hit = np.zeros_like(before)
                for v in vals:
                    hit |= np.asarray(self._binop("==", sel, self._eval(v, fr)), dtype=bool)
                hit &= before & ~matched  # the first matching clause wins
                matched |= hit
                if hit.any():
                    fr.active = hit
                    self._block(body, fr)
                    after |= fr.active
            rest = before & ~matched
            if s[3] is not None and rest.any():
                fr.active = rest
                self._block(s[3], fr)
                after |= fr.active
            elif s[3] is None:
                after |= rest
            fr.active = after
        elif kind in ("let", "var"):
            if kind == "var":
                raise NotImplementedError("SIMT executor: function-scope `var` (none of the step shaders has one)")
            v = self._eval(s[3], fr)
            if _is_abstract(v):
                v = f32(v) if isinstance(v, float) else i32(v)
            fr.scopes[-1][s[1]] = v
        elif kind == "assign":
            op, lhs, rhs = s[1], s[2], s[3]
            val = self._eval(rhs, fr)
            if op != "=":
                val = self._binop(op[0], self._eval(lhs, fr), val)
            if lhs[0] != "index":
                raise NotImplementedError("SIMT executor: assignment to a local")
            arr = self._eval(lhs[1], fr)
            idx = self._eval(lhs[2], fr)
            self._store(arr[1], idx, val, fr.active)
        elif kind == "expr":
            self._eval(s[1], fr)
        else:
            raise NotImplementedError(f"SIMT executor: statement `{kind}`")

    # -- memory ----------------------------------------------------------------------------------------------
    def _is_own(self, idx, active):
        return idx is self.gid or (isinstance(idx, np.ndarray) and np.array_equal(idx[active], self.gid[active]))

    def _load(self, name, idx, active):
        data = self.d.arrays[name]
        m = len(data)
        if OOB_POLICY == "clamp" and m > 0:
            # the other thing a WebGPU backend may do with an out-of-range read (naga's `Restrict` policy, Tint's
            # robustness transform): min(u32(index), length - 1).  Sensitivity study only (tests/test_wgsl_pin.py);
            # the oracle's and the product's semantics are "reads return 0" (SURVEY.md section 8).
            if not isinstance(idx, np.ndarray):
                return data[min(int(idx) & 0xFFFFFFFF, m - 1)]
            if not self._is_own(idx, active):
                self.d.foreign.add(name)
                if name in self.d.stored:
                    raise self.d.conflict(name)
            return data[np.minimum(idx.astype(np.int64) & 0xFFFFFFFF, m - 1)]
        if idx is self.gid:
            if m >= self.hi:
                return data[self.lo:self.hi].copy()
            out = np.zeros(self.n, dtype=data.dtype)
            if m > self.lo:
                out[:m - self.lo] = data[self.lo:]
            return out
        if not self._is_own(idx, active):
            self.d.foreign.add(name)
            if name in self.d.stored:
                raise self.d.conflict(name)
        if not isinstance(idx, np.ndarray):
            i = int(idx)
            return data[i] if 0 <= i < m else data.dtype.type(0)
        idx64 = idx.astype(np.int64)
        ok = (idx64 >= 0) & (idx64 < m)
        out = data[np.where(ok, idx64, 0)]
        out[~ok] = 0
        return out

    def _store(self, name, idx, val, active):
        data = self.d.arrays[name]
        if not self._is_own(idx, active):
            raise self.d.conflict(name)
        self.d.stored.add(name)
        if name in self.d.foreign:
            raise self.d.conflict(name)
        k = max(0, min(len(data), self.hi) - self.lo)  # lanes whose own element exists: the rest is dropped
        if isinstance(val, Vec3):
            if data.ndim != 2 or data.shape[1] != 3:
                raise TypeError(f"WGSL: storing a vec3 into `{name}`")
            for c, comp in enumerate(val.c):
                if _is_abstract(comp):
                    comp = data.dtype.type(comp)
                if isinstance(comp, np.ndarray):
                    if comp.dtype != data.dtype:
                        raise TypeError(f"WGSL: storing vec3<{comp.dtype}> into `{name}`")
                    np.copyto(data[self.lo:self.lo + k, c], comp[:k], where=active[:k])
                else:
                    data[self.lo:self.lo + k, c][active[:k]] = comp
            return
        if _is_abstract(val):
            val = data.dtype.type(val)
        if isinstance(val, np.ndarray):
            if val.dtype != data.dtype:
                raise TypeError(f"WGSL: storing {val.dtype} into array<{data.dtype}> `{name}`")
            np.copyto(data[self.lo:self.lo + k], val[:k], where=active[:k])
        else:
            if np.dtype(type(val)) != data.dtype:
                raise TypeError(f"WGSL: storing {type(val)} into array<{data.dtype}> `{name}`")
            data[self.lo:self.lo + k][active[:k]] = val

    # -- expressions -----------------------------------------------------------------------------------------
    def _lookup(self, name, fr):
        for sc in reversed(fr.scopes):
            if name in sc:
                return sc[name]
        raise NameError(f"WGSL: undefined `{name}`")

    def _eval(self, e, fr):
        kind = e[0]
        if kind == "lit":
            return _CONCRETE.get(e[2], lambda x: x)(e[1])
        if kind == "var":
            return self._lookup(e[1], fr)
        if kind == "member":
            base = self._eval(e[1], fr)
            if isinstance(base, dict):
                return base[e[2]]
            raise TypeError(f"WGSL: member .{e[2]} of {type(base)}")
        if kind == "index":
            arr = self._eval(e[1], fr)
            if not (isinstance(arr, tuple) and arr[0] == "array"):
                raise TypeError("WGSL: indexing a non-array")
            return self._load(arr[1], self._eval(e[2], fr), fr.active)
        if kind == "neg":
            return -self._eval(e[1], fr)
        if kind == "not":
            return np.logical_not(self._eval(e[1], fr))
        if kind == "bin":
            op = e[1]
            a, b = self._eval(e[2], fr), self._eval(e[3], fr)  # no side effects in expressions: no short circuit needed
            if op == "&&":
                return np.logical_and(a, b)
            if op == "||":
                return np.logical_or(a, b)
            return self._binop(op, a, b)
        if kind == "call":
            return self._call(e[1], [self._eval(a, fr) for a in e[2]], fr)
        raise NotImplementedError(kind)

    @staticmethod
    def _binop(op, a, b):
        if isinstance(a, Vec3) or isinstance(b, Vec3):  # component-wise, scalars are splat
            ac = a.c if isinstance(a, Vec3) else [a] * 3
            bc = b.c if isinstance(b, Vec3) else [b] * 3
            if op not in "+-*/":
                raise NotImplementedError(f"vec3 operator {op}")
            return Vec3([_Lanes._binop(op, x, y) for x, y in zip(ac, bc)])
        a, b = _unify(a, b)
        if not _is_abstract(a):
            ta = a.dtype if isinstance(a, np.ndarray) else np.dtype(type(a))
            tb = b.dtype if isinstance(b, np.ndarray) else np.dtype(type(b))
            if ta != tb:
                raise TypeError(f"WGSL: operands of `{op}` have types {ta} and {tb}")
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            if _is_int(a):
                if _is_abstract(a):
                    return a if b == 0 else a // b
                zero = b == 0
                q = np.floor_divide(a, np.where(zero, 1, b).astype(np.asarray(a).dtype))
                return np.where(zero, a, q).astype(np.asarray(a).dtype) if np.ndim(q) else (a if zero else q)
            return a / b
        if op == "%":
            if _is_int(a):
                if _is_abstract(a):
                    return 0 if b == 0 else a % b
                zero = b == 0
                r = np.remainder(a, np.where(zero, 1, b).astype(np.asarray(a).dtype))
                return np.where(zero, 0, r).astype(np.asarray(a).dtype) if np.ndim(r) else (type(a)(0) if zero else r)
            return np.fmod(a, b)
        if op == "==":
            return a == b
        if op == "!=":
            return a != b
        if op == "<":
            return a < b
        if op == ">":
            return a > b
        if op == "<=":
            return a <= b
        if op == ">=":
            return a >= b
        raise NotImplementedError(op)

    def _call(self, name, args, fr):
        if name in self.sh.funcs:
            fn = self.sh.funcs[name]
            scope = {}
            for (pname, pty, _), a in zip(fn["params"], args):
                if _is_abstract(a):
                    a = _CONCRETE[pty[0]](a)
                scope[pname] = a
            inner = _Frame([self.d.env, scope], fr.active.copy())
            self._block(fn["body"], inner)
            return inner.ret
        if name == "vec3":
            return Vec3(args * 3 if len(args) == 1 else args)
        if name == "i32":
            a = args[0]
            if isinstance(a, np.ndarray) and a.dtype.kind == "f":  # truncate toward zero, saturate, NaN -> 0
                t = np.clip(np.trunc(np.where(np.isnan(a), 0, a)), -2147483648.0, 2147483520.0)
                return t.astype(i32)
            if isinstance(a, np.ndarray):
                return a.astype(i32)
            return i32(int(a))
        if name == "f32":
            a = args[0]
            return a.astype(f32) if isinstance(a, np.ndarray) else f32(a)
        if name == "u32":
            a = args[0]
            return a.astype(u32) if isinstance(a, np.ndarray) else u32(int(a) & 0xffffffff)
        if name == "sqrt":
            a = args[0]
            return np.sqrt(f32(a)) if _is_abstract(a) else np.sqrt(a)
        if name == "floor":
            a = args[0]
            return np.floor(f32(a)) if _is_abstract(a) else np.floor(a)
        if name == "abs":
            return np.abs(args[0])
        if name in ("min", "max"):
            a, b = _unify(args[0], args[1])
            return (np.minimum if name == "min" else np.maximum)(a, b)
        if name == "clamp":
            x, lo, hi = args
            if _is_abstract(lo):
                lo = _concretise(lo, x)
            if _is_abstract(hi):
                hi = _concretise(hi, x)
            return np.minimum(np.maximum(x, lo), hi)  # WGSL: clamp(e, low, high) = min(max(e, low), high)
        raise NotImplementedError(f"SIMT executor: function `{name}`")


class WgslLBMVec(WgslLBM):
    """WgslLBM (the reference's host-side dispatch order, lbm.rs) with the step and summary shaders executed by the
    SIMT executor, colour maps included; only barrier_draw.wgsl (a scatter) keeps the scalar interpreter."""


    def __init__(self, omega, x, y, inflow_ux=0.1, root=None, threads=1):
        from . import wgsl_interp
        root = root or wgsl_interp.SHADER_ROOT
        super().__init__(omega, x, y, inflow_ux=inflow_ux, root=root)
        rel = {"pre_corner": "pre_collision/corner_pre_collision.wgsl",
               "pre_cardinal": "pre_collision/cardinal_pre_collision.wgsl",
               "col_cardinal": "collision/cardinal_collision.wgsl", "col_corner": "collision/corner_collision.wgsl",
               "ne_sw": "stream/ne_sw_stream.wgsl", "nw_se": "stream/se_nw_stream.wgsl",
               "n_s": "stream/n_s_stream.wgsl", "e_w": "stream/e_w_stream.wgsl", "ux": "summary_stats/ux.wgsl",
               "uy": "summary_stats/uy.wgsl", "rho": "summary_stats/rho.wgsl", "speed": "summary_stats/speed.wgsl",
               "curl": "summary_stats/curl.wgsl", "inferno": "color_map/inferno.wgsl",
               "viridis": "color_map/viridis.wgsl", "jet": "color_map/jet.wgsl"}
        self.vsh = {k: VecShader(os.path.join(root, v)) for k, v in rel.items()}
        self.threads = int(threads)
        self.pool = ThreadPoolExecutor(self.threads) if self.threads > 1 else None

    def color_map(self, name=None):
        name = name or self.color_map_name
        self.vsh[name].dispatch(self.work_groups, {(0, 0): self.colors, (1, 0): self.output, (2, 0): self.barrier,
                                                   (3, 0): u32(self.n)}, threads=self.threads, pool=self.pool)

    def _run(self, name, bindings):
        if name in self.vsh:
            self.vsh[name].dispatch(self.work_groups, bindings, threads=self.threads, pool=self.pool)
        else:
            self.sh[name].dispatch(self.work_groups, bindings)
