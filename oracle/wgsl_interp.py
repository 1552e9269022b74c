"""TEST INFRASTRUCTURE — an interpreter for the subset of WGSL the reference's compute shaders use, and a
driver that dispatches them in the reference's order.

Why: the reference (lbm-wgpu) cannot be built or run in this environment (Rust -> wasm32 + browser WebGPU, no
cargo / node / Vulkan), so the C oracle (oracle/lbm_oracle.c) is a hand restatement.  What *is* available is the
text of the reference's shaders.  This module parses that text as it lies under
/root/reference/lbm-wgpu/src/rewritten_shaders and evaluates every invocation with numpy fp32 / u32 scalars, so
the arithmetic (association, operation order, guards, index helpers) comes from the reference's own source and
not from a transcription.  tests/golden/make_wgsl_golden.py uses it to generate golden vectors that pin the C
oracle and the CUDA path; tests/test_wgsl_pin.py re-runs it live wherever /root/reference exists.

Semantics assumed where WGSL / the WebGPU backend leave room (the same choices as the oracle's normative block,
SURVEY.md section 8): every binary fp32 operation individually rounded (no FMA contraction), IEEE division and
sqrt, u32 arithmetic wraps, out-of-range array reads return 0, out-of-range writes are dropped; constant
expressions of abstract literals (`1.0/36.0`) are evaluated in f64 and then rounded to f32 (for the three
constants in these shaders that equals the fp32 quotient, asserted in the tests).

The host side (which buffer is bound where, pass order, ping-pong index) follows lbm.rs and is cited per method.
Only tests/ and the golden generator may import this; the product never does.
"""
import os
import re

import numpy as np

SHADER_ROOT = "/root/reference/lbm-wgpu/src/rewritten_shaders"

f32 = np.float32
u32 = np.uint32
i32 = np.int32

# ---------------------------------------------------------------------------------------------------------------
# tokenizer
# ---------------------------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*)
  | (?P<float>(\d+\.\d*|\.\d+)([eE][+-]?\d+)?f?|\d+[eE][+-]?\d+f?|\d+f)
  | (?P<uint>\d+u)
  | (?P<int>\d+i?)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>\+=|-=|\*=|/=|==|!=|<=|>=|&&|\|\||->|[-+*/%<>=!(){}\[\];:,.@&|])
""", re.X)


def tokenize(src):
    out, pos = [], 0
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"WGSL: cannot tokenize at {src[pos:pos + 30]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        out.append((kind, m.group(kind)))
    out.append(("eof", ""))
    return out


# ---------------------------------------------------------------------------------------------------------------
# parser -> tuples
# ---------------------------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, src):
        self.t = tokenize(src)
        self.p = 0

    def peek(self, k=0):
        return self.t[self.p + k]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "id"):
            self.p += 1
            return True
        return False

    def expect(self, val):
        tok = self.next()
        if tok[1] != val:
            raise SyntaxError(f"WGSL: expected {val!r}, got {tok!r}")

    def ident(self):
        tok = self.next()
        if tok[0] != "id":
            raise SyntaxError(f"WGSL: expected identifier, got {tok!r}")
        return tok[1]

    # -- types: u32 | f32 | i32 | vec3<f32> | array<T> | Name
    def type_(self):
        name = self.ident()
        if self.accept("<"):
            inner = self.type_()
            self.expect(">")
            return (name, inner)
        return (name, None)

    def attributes(self):
        attrs = {}
        while self.accept("@"):
            name = self.ident()
            args = []
            if self.accept("("):
                while not self.accept(")"):
                    args.append(self.next()[1])
                    self.accept(",")
            attrs[name] = args
        return attrs

    def module(self):
        structs, globals_, funcs = {}, {}, {}
        while self.peek()[0] != "eof":
            attrs = self.attributes()
            if self.accept("struct"):
                name = self.ident()
                self.expect("{")
                fields = []
                while not self.accept("}"):
                    self.attributes()
                    fname = self.ident()
                    self.expect(":")
                    fields.append((fname, self.type_()))
                    self.accept(",")
                self.accept(";")
                structs[name] = fields
            elif self.accept("var"):
                space = None
                if self.accept("<"):
                    space = self.ident()
                    while not self.accept(">"):
                        self.next()
                name = self.ident()
                self.expect(":")
                ty = self.type_()
                self.expect(";")
                globals_[name] = {"group": int(attrs["group"][0]), "binding": int(attrs["binding"][0]),
                                  "space": space, "type": ty}
            elif self.accept("fn"):
                name = self.ident()
                self.expect("(")
                params = []
                while not self.accept(")"):
                    pattrs = self.attributes()
                    pname = self.ident()
                    self.expect(":")
                    params.append((pname, self.type_(), pattrs))
                    self.accept(",")
                ret = None
                if self.accept("->"):
                    self.attributes()
                    ret = self.type_()
                body = self.block()
                funcs[name] = {"params": params, "ret": ret, "body": body, "attrs": attrs}
            else:
                raise SyntaxError(f"WGSL: unexpected token {self.peek()!r} at module scope")
        return structs, globals_, funcs

    def block(self):
        self.expect("{")
        stmts = []
        while not self.accept("}"):
            stmts.append(self.statement())
        return stmts

    def statement(self):
        if self.peek()[1] == "{" and self.peek()[0] == "op":
            return ("block", self.block())
        if self.accept("return"):
            e = None
            if not self.accept(";"):
                e = self.expr()
                self.expect(";")
            return ("return", e)
        if self.accept("if"):
            cond = self.expr()
            then = self.block()
            other = None
            if self.accept("else"):
                other = [self.statement()] if self.peek()[1] == "if" else self.block()
            return ("if", cond, then, other)
        if self.accept("switch"):
            sel = self.expr()
            self.expect("{")
            cases, default = [], None
            while not self.accept("}"):
                if self.accept("default"):
                    self.accept(":")
                    default = self.block()
                else:
                    self.expect("case")
                    vals = [self.expr()]
                    while self.accept(","):
                        vals.append(self.expr())
                    self.accept(":")
                    cases.append((vals, self.block()))
            return ("switch", sel, cases, default)
        if self.peek()[1] in ("let", "var") and self.peek()[0] == "id":
            kind = self.next()[1]
            name = self.ident()
            ty = None
            if self.accept(":"):
                ty = self.type_()
            init = None
            if self.accept("="):
                init = self.expr()
            self.expect(";")
            return (kind, name, ty, init)
        lhs = self.expr()
        for op in ("=", "+=", "-=", "*=", "/="):
            if self.accept(op):
                rhs = self.expr()
                self.expect(";")
                return ("assign", op, lhs, rhs)
        self.expect(";")
        return ("expr", lhs)

    # precedence climbing: || < && < comparison < additive < multiplicative < unary < postfix
    def expr(self):
        return self.binary(0)

    LEVELS = [("||",), ("&&",), ("==", "!=", "<", ">", "<=", ">="), ("+", "-"), ("*", "/", "%")]

    def binary(self, level):
        if level == len(self.LEVELS):
            return self.unary()
        lhs = self.binary(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in self.LEVELS[level]:
            op = self.next()[1]
            rhs = self.binary(level + 1)
            lhs = ("bin", op, lhs, rhs)  # left-associative, as WGSL
        return lhs

    def unary(self):
        if self.accept("-"):
            return ("neg", self.unary())
        if self.accept("!"):
            return ("not", self.unary())
        return self.postfix()

    def postfix(self):
        kind, val = self.next()
        if kind == "float":
            e = ("lit", float(val.rstrip("f")), "abstract-float" if not val.endswith("f") else "f32")
        elif kind == "uint":
            e = ("lit", int(val[:-1]), "u32")
        elif kind == "int":
            e = ("lit", int(val.rstrip("i")), "i32" if val.endswith("i") else "abstract-int")
        elif kind == "op" and val == "(":
            e = self.expr()
            self.expect(")")
        elif kind == "id":
            if self.peek()[1] == "<" and val in ("vec2", "vec3", "vec4", "array"):
                self.next()
                self.type_()
                self.expect(">")
            if self.accept("("):
                args = []
                while not self.accept(")"):
                    args.append(self.expr())
                    self.accept(",")
                e = ("call", val, args)
            else:
                e = ("var", val)
        else:
            raise SyntaxError(f"WGSL: unexpected token {(kind, val)!r} in expression")
        while True:
            if self.accept("["):
                idx = self.expr()
                self.expect("]")
                e = ("index", e, idx)
            elif self.accept("."):
                e = ("member", e, self.ident())
            else:
                return e


# ---------------------------------------------------------------------------------------------------------------
# evaluation
# ---------------------------------------------------------------------------------------------------------------
class _Return(Exception):
    def __init__(self, value):
        self.value = value


def _is_abstract(v):
    return isinstance(v, (int, float)) and not isinstance(v, (bool, np.generic))


def _concretise(v, like):
    """abstract literal -> the concrete type of the other operand"""
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


class StorageArray:
    """array<T> in a storage buffer with WebGPU robust-access semantics as the oracle defines them"""

    def __init__(self, data):
        self.data = data

    def load(self, i):
        i = int(i)
        if 0 <= i < len(self.data):
            return self.data[i]
        if self.data.dtype.names:
            return np.zeros((), dtype=self.data.dtype)[()]
        return self.data.dtype.type(0)

    def store(self, i, v):
        i = int(i)
        if 0 <= i < len(self.data):
            self.data[i] = v


class Shader:
    def __init__(self, path_or_src, is_source=False):
        self.path = None if is_source else path_or_src
        src = path_or_src if is_source else open(path_or_src).read()
        self.structs, self.globals, self.funcs = Parser(src).module()
        if "main" not in self.funcs or "compute" not in self.funcs["main"]["attrs"]:
            raise SyntaxError("WGSL: no @compute fn main")
        self.workgroup_size = int(self.funcs["main"]["attrs"]["workgroup_size"][0])

    # bindings: {(group, binding): numpy array | numpy scalar | dict (uniform struct)}
    def dispatch(self, workgroups, bindings):
        env_globals = {}
        for name, g in self.globals.items():
            key = (g["group"], g["binding"])
            if key not in bindings:
                raise KeyError(f"WGSL: nothing bound at @group({key[0]}) @binding({key[1]}) for `{name}`")
            b = bindings[key]
            is_array = g["type"][0] == "array"
            if is_array != isinstance(b, np.ndarray):
                raise TypeError(f"WGSL: binding for `{name}` has the wrong shape")
            env_globals[name] = StorageArray(b) if is_array else b
        with np.errstate(all="ignore"):
            for gid in range(workgroups * self.workgroup_size):
                self._invoke(env_globals, gid)

    def _invoke(self, env_globals, gid):
        fn = self.funcs["main"]
        scope = {}
        for pname, _, pattrs in fn["params"]:
            if pattrs.get("builtin") == ["global_invocation_id"]:
                scope[pname] = {"x": u32(gid), "y": u32(0), "z": u32(0)}
            else:
                raise NotImplementedError(f"WGSL: builtin {pattrs}")
        try:
            self._exec_block(fn["body"], [env_globals, scope])
        except _Return:
            pass

    # -- statements
    def _exec_block(self, stmts, scopes):
        scopes = scopes + [{}]
        for s in stmts:
            self._exec(s, scopes)

    def _exec(self, s, scopes):
        kind = s[0]
        if kind == "return":
            raise _Return(None if s[1] is None else self._eval(s[1], scopes))
        if kind == "block":
            self._exec_block(s[1], scopes)
        elif kind == "if":
            if bool(self._eval(s[1], scopes)):
                self._exec_block(s[2], scopes)
            elif s[3] is not None:
                self._exec_block(s[3], scopes)
        elif kind == "switch":
            sel = self._eval(s[1], scopes)
            for vals, body in s[2]:
                if any(int(self._eval(v, scopes)) == int(sel) for v in vals):
                    self._exec_block(body, scopes)
                    return
            if s[3] is not None:
                self._exec_block(s[3], scopes)
        elif kind in ("let", "var"):
            v = self._eval(s[3], scopes) if s[3] is not None else None
            if _is_abstract(v):  # a let of an abstract value takes the default concrete type
                v = f32(v) if isinstance(v, float) else i32(v)
            scopes[-1][s[1]] = v
        elif kind == "assign":
            op, lhs, rhs = s[1], s[2], s[3]
            val = self._eval(rhs, scopes)
            if op != "=":
                val = self._binop(op[0], self._eval(lhs, scopes), val)  # a += e  ==  a = a + (e)
            self._store(lhs, val, scopes)
        elif kind == "expr":
            self._eval(s[1], scopes)
        else:
            raise NotImplementedError(kind)

    def _store(self, lhs, val, scopes):
        if lhs[0] == "index":
            arr = self._eval(lhs[1], scopes)
            idx = self._eval(lhs[2], scopes)
            if not isinstance(arr, StorageArray):
                raise TypeError("WGSL: indexed store into a non-array")
            if _is_abstract(val):
                val = arr.data.dtype.type(val)
            arr.store(idx, val)
        elif lhs[0] == "var":
            for sc in reversed(scopes):
                if lhs[1] in sc:
                    sc[lhs[1]] = val
                    return
            raise NameError(lhs[1])
        else:
            raise NotImplementedError(f"WGSL: store to {lhs[0]}")

    # -- expressions
    def _lookup(self, name, scopes):
        for sc in reversed(scopes):
            if name in sc:
                return sc[name]
        raise NameError(f"WGSL: undefined `{name}`")

    def _eval(self, e, scopes):
        kind = e[0]
        if kind == "lit":
            v, ty = e[1], e[2]
            return {"u32": u32, "i32": i32, "f32": f32}.get(ty, lambda x: x)(v)
        if kind == "var":
            return self._lookup(e[1], scopes)
        if kind == "member":
            base = self._eval(e[1], scopes)
            if isinstance(base, dict):
                return base[e[2]]
            if isinstance(base, np.void):
                return base[e[2]]
            if isinstance(base, np.ndarray) and e[2] in "xyz":
                return base["xyz".index(e[2])]
            raise TypeError(f"WGSL: member .{e[2]} of {type(base)}")
        if kind == "index":
            arr = self._eval(e[1], scopes)
            idx = self._eval(e[2], scopes)
            return arr.load(idx)
        if kind == "neg":
            v = self._eval(e[1], scopes)
            return -v
        if kind == "not":
            return not bool(self._eval(e[1], scopes))
        if kind == "bin":
            op = e[1]
            if op == "&&":
                return bool(self._eval(e[2], scopes)) and bool(self._eval(e[3], scopes))
            if op == "||":
                return bool(self._eval(e[2], scopes)) or bool(self._eval(e[3], scopes))
            return self._binop(op, self._eval(e[2], scopes), self._eval(e[3], scopes))
        if kind == "call":
            return self._call(e[1], [self._eval(a, scopes) for a in e[2]], scopes)
        raise NotImplementedError(kind)

    @staticmethod
    def _binop(op, a, b):
        a, b = _unify(a, b)
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            if isinstance(a, (np.integer, int)) and not isinstance(a, bool):
                if int(b) == 0:
                    return a  # WGSL: x / 0 == x for integers
                return type(a)(int(a) // int(b)) if isinstance(a, np.generic) else a // b
            return a / b
        if op == "%":
            if isinstance(a, (np.integer, int)):
                if int(b) == 0:
                    return type(a)(0) if isinstance(a, np.generic) else 0
                return type(a)(int(a) % int(b)) if isinstance(a, np.generic) else a % b
            return np.fmod(a, b)
        if op == "==":
            return bool(a == b)
        if op == "!=":
            return bool(a != b)
        if op == "<":
            return bool(a < b)
        if op == ">":
            return bool(a > b)
        if op == "<=":
            return bool(a <= b)
        if op == ">=":
            return bool(a >= b)
        raise NotImplementedError(op)

    def _call(self, name, args, scopes):
        if name in self.funcs:
            fn = self.funcs[name]
            scope = {}
            for (pname, pty, _), a in zip(fn["params"], args):
                if _is_abstract(a):
                    a = {"u32": u32, "i32": i32, "f32": f32}[pty[0]](a)
                scope[pname] = a
            try:
                self._exec_block(fn["body"], [scopes[0], scope])
            except _Return as r:
                return r.value
            return None
        if name in ("vec3", "vec2", "vec4"):
            if len(args) == 1:
                args = args * int(name[3])
            return np.array([f32(a) for a in args], dtype=f32)
        if name == "f32":
            return f32(args[0])
        if name == "u32":
            return u32(int(args[0]) & 0xffffffff)
        if name == "i32":
            a = args[0]
            if isinstance(a, (np.floating, float)):
                a = np.trunc(a)
                return i32(int(min(max(a, -2147483648.0), 2147483520.0)))
            return i32(int(a))
        if name == "sqrt":
            a = args[0]
            return np.sqrt(f32(a)) if _is_abstract(a) else np.sqrt(a)
        if name == "floor":
            a = args[0]
            return np.floor(f32(a)) if _is_abstract(a) else np.floor(a)
        if name == "abs":
            return abs(args[0])
        if name in ("min", "max"):
            a, b = _unify(args[0], args[1])
            return (np.minimum if name == "min" else np.maximum)(a, b)
        if name == "clamp":
            x = args[0]
            lo, hi = args[1], args[2]
            if _is_abstract(lo):
                lo = _concretise(lo, x)
            if _is_abstract(hi):
                hi = _concretise(hi, x)
            return np.minimum(np.maximum(x, lo), hi)  # WGSL: clamp(e, low, high) = min(max(e, low), high)
        raise NotImplementedError(f"WGSL: function `{name}`")


# ---------------------------------------------------------------------------------------------------------------
# the reference's host side: which buffer is bound where, in which order the passes run
# ---------------------------------------------------------------------------------------------------------------
NW, N, NE, W, REST, E, SW, S, SE = range(9)  # data_buffers[b][k], lbm.rs:632-640


def shaders_available(root=SHADER_ROOT):
    return os.path.isdir(root)


class WgslLBM:
    """`pub struct LBM` (lbm.rs:32-98) with every compute pass executed by interpreting the reference's WGSL.
    Same method names as the C oracle's Python wrapper (oracle/lbm_oracle.py) where they overlap."""

    # lbm.rs:822-900, 969: which shader file each pipeline is built from
    SHADERS = {
        "pre_corner": "pre_collision/corner_pre_collision.wgsl",
        "pre_cardinal": "pre_collision/cardinal_pre_collision.wgsl",
        "col_cardinal": "collision/cardinal_collision.wgsl",
        "col_corner": "collision/corner_collision.wgsl",
        "ne_sw": "stream/ne_sw_stream.wgsl",
        "nw_se": "stream/se_nw_stream.wgsl",
        "n_s": "stream/n_s_stream.wgsl",
        "e_w": "stream/e_w_stream.wgsl",
        "ux": "summary_stats/ux.wgsl",
        "uy": "summary_stats/uy.wgsl",
        "rho": "summary_stats/rho.wgsl",
        "speed": "summary_stats/speed.wgsl",
        "curl": "summary_stats/curl.wgsl",
        "draw": "update_barrier/barrier_draw.wgsl",
        "inferno": "color_map/inferno.wgsl",
        "viridis": "color_map/viridis.wgsl",
        "jet": "color_map/jet.wgsl",
    }

    def _load_shaders(self, root):
        self.sh = {name: Shader(os.path.join(root, rel)) for name, rel in self.SHADERS.items()}

    def __init__(self, omega, x, y, inflow_ux=0.1, root=SHADER_ROOT):
        self.x, self.y = int(x), int(y)
        self.n = self.x * self.y
        self._load_shaders(root)
        self.omega = f32(omega)
        # lbm.rs:739-744: both buffer sets from the same initial data
        init = self.set_equil(f32(inflow_ux), f32(0.0), f32(1.0))
        self.data = [[np.full(self.n, init[k], dtype=f32) for k in range(9)] for _ in range(2)]
        self.barrier = self.init_barrier()
        # density_bg (lbm.rs:780-782) and output_bg (:783-785) start as zeros
        self.ux = np.zeros(self.n, f32)
        self.uy = np.zeros(self.n, f32)
        self.rho = np.zeros(self.n, f32)
        self.output = np.zeros(self.n, f32)
        self.colors = np.zeros((self.n, 3), f32)
        self.compute_step = 0
        # dispatch size: lbm.rs work_group_size = ceil(x*y / 256)
        self.work_groups = (self.n + 255) // 256
        self.summary_stat = "curl"
        self.color_map_name = "jet"

    # -- lbm.rs:611-643, the host-side fp32 evaluation of the initial equilibrium (same op order)
    @staticmethod
    def set_equil(ux, uy, rho):
        ux, uy, rho = f32(ux), f32(uy), f32(rho)
        ux_2 = ux * ux
        uy_2 = uy * uy
        u_dot_product = ux_2 + uy_2
        u_sum_sq_pos = u_dot_product + f32(2.0) * (ux * uy)
        u_sum_sq_neg = u_dot_product - f32(2.0) * (ux * uy)
        ux = ux * f32(3.0)
        uy = uy * f32(3.0)
        ux_2 = ux_2 * f32(4.5)
        uy_2 = uy_2 * f32(4.5)
        u_dot_product = u_dot_product * f32(1.5)
        u_sum_sq_neg = u_sum_sq_neg * f32(4.5)
        u_sum_sq_pos = u_sum_sq_pos * f32(4.5)
        rho_ninth = rho / f32(9.0)
        rho_36th = rho / f32(36.0)
        one = f32(1.0)
        out = [None] * 9
        out[NW] = rho_36th * (one - ux + uy + u_sum_sq_neg - u_dot_product)
        out[N] = rho_ninth * (one + uy + uy_2 - u_dot_product)
        out[NE] = rho_36th * (one + ux + uy + u_sum_sq_pos - u_dot_product)
        out[W] = rho_ninth * (one - ux + ux_2 - u_dot_product)
        out[REST] = f32(4.0) * rho_ninth * (one - u_dot_product)
        out[E] = rho_ninth * (one + ux + ux_2 - u_dot_product)
        out[SW] = rho_36th * (one - ux - uy + u_sum_sq_pos - u_dot_product)
        out[S] = rho_ninth * (one - uy - uy_2 - u_dot_product)
        out[SE] = rho_36th * (one + ux - uy + u_sum_sq_neg - u_dot_product)
        return out

    # -- lbm.rs:595-605
    def init_barrier(self):
        b = np.zeros(self.n, dtype=u32)
        b[: self.x] = 1
        b[(self.y - 1) * self.x:] = 1
        return b

    # -- bind groups (lbm.rs:760-778, 786-790)
    def _pair(self, which, buf):
        ka, kb = {"ne_sw": (NE, SW), "nw_se": (NW, SE), "n_s": (N, S), "e_w": (E, W)}[which]
        return self.data[buf][ka], self.data[buf][kb]

    def _density(self, group):
        return {(group, 0): self.ux, (group, 1): self.uy, (group, 2): self.rho}

    def _dimensions(self):
        # create_dimension_bg: row = x, col = y, total = x*y
        return {"row": u32(self.x), "col": u32(self.y), "total": u32(self.n)}

    def _run(self, name, bindings):
        self.sh[name].dispatch(self.work_groups, bindings)

    # -- lbm.rs:1174-1212
    def collide(self):
        c = self.compute_step % 2
        size = u32(self.n)
        rest = self.data[0][REST]  # collide_bg is built once from data_buffers[0][4] (lbm.rs:775-778)
        ne, sw = self._pair("ne_sw", c)
        nw, se = self._pair("nw_se", c)
        n, s = self._pair("n_s", c)
        e, w = self._pair("e_w", c)
        self._run("pre_corner", {(0, 0): ne, (0, 1): sw, (1, 0): nw, (1, 1): se, **self._density(2), (3, 0): size})
        self._run("pre_cardinal", {(0, 0): n, (0, 1): s, (1, 0): e, (1, 1): w, **self._density(2), (3, 0): size})
        self._run("col_corner", {(0, 0): ne, (0, 1): sw, (1, 0): nw, (1, 1): se, **self._density(2),
                                 (3, 0): size, (3, 1): self.omega, (3, 2): rest})
        self._run("col_cardinal", {(0, 0): n, (0, 1): s, (1, 0): e, (1, 1): w, **self._density(2),
                                   (3, 0): size, (3, 1): self.omega, (3, 2): rest})

    # -- lbm.rs:1127-1134, 1214-1252: e_w, n_s, nw_se, ne_sw
    def stream(self):
        c, d = self.compute_step % 2, (self.compute_step + 1) % 2
        for which in ("e_w", "n_s", "nw_se", "ne_sw"):
            a, b = self._pair(which, c)
            pa, pb = self._pair(which, d)
            self._run(which, {(0, 0): self._dimensions(), (1, 0): a, (1, 1): b, (2, 0): pa, (2, 1): pb,
                              (3, 0): self.barrier})

    # -- lbm.rs:1112-1116
    def step(self):
        self.collide()
        self.stream()
        self.compute_step += 1

    # -- lbm.rs:1051-1059, 1254-1297
    def calculate_summary(self):
        self._run(self.summary_stat, {(0, 0): self._dimensions(), **self._density(1), (2, 0): self.output})

    # -- lbm.rs:1065-1074 minus colour/render
    def iterate(self, n):
        for _ in range(int(n)):
            self.step()
        self.calculate_summary()

    def set_summary(self, name):
        self.summary_stat = name

    # -- lbm.rs:1299-1335; NB the reference declares array<vec3<f32>> (16-byte stride) over a 12 B/cell buffer:
    #    a latent bug that is not reproduced here; colours are kept as 3 floats per cell
    def color_map(self, name=None):
        name = name or self.color_map_name
        colors = _Vec3Array(self.colors)
        self.sh[name].dispatch_objects(self.work_groups, {(0, 0): colors, (1, 0): self.output,
                                                          (2, 0): self.barrier, (3, 0): u32(self.n)})

    # -- lbm.rs:1337-1356 + barrier_draw.wgsl: pairs (location, value)
    def draw_points(self, pairs):
        pairs = np.asarray(pairs, dtype=np.uint32).reshape(-1, 2)
        if len(pairs) == 0:
            return
        upd = np.zeros(len(pairs), dtype=[("location", np.uint32), ("value", np.uint32)])
        upd["location"] = pairs[:, 0]
        upd["value"] = pairs[:, 1]
        # draw_barrier_updates dispatches one workgroup (size 1) per update
        self.sh["draw"].dispatch(len(pairs), {(0, 0): u32(len(pairs)), (0, 1): upd, (1, 0): self.barrier})

    # -- lbm.rs:1358-1360
    def update_omega_buffer(self, omega):
        self.omega = f32(omega)

    # -- lbm.rs:1362-1365
    def reset_barrier(self):
        self.barrier[:] = self.init_barrier()

    # -- lbm.rs:1076-1102: re-init, step = 0, the two pre-collision passes only
    def custom_speed(self, ux):
        init = self.set_equil(f32(ux), f32(0.0), f32(1.0))
        for b in range(2):
            for k in range(9):
                self.data[b][k][:] = init[k]
        self.compute_step = 0
        c = 0
        size = u32(self.n)
        ne, sw = self._pair("ne_sw", c)
        nw, se = self._pair("nw_se", c)
        n, s = self._pair("n_s", c)
        e, w = self._pair("e_w", c)
        self._run("pre_corner", {(0, 0): ne, (0, 1): sw, (1, 0): nw, (1, 1): se, **self._density(2), (3, 0): size})
        self._run("pre_cardinal", {(0, 0): n, (0, 1): s, (1, 0): e, (1, 1): w, **self._density(2), (3, 0): size})

    def reset_to_equilibrium(self):
        self.custom_speed(f32(0.1))

    # -- lbm.rs:1482-1515: set_equil(0,0,1) plus one population set to 4.0 at a fixed cell; no pre-collision
    def single_cell(self, index):
        x, y = self.x, self.y
        init = self.set_equil(f32(0.0), f32(0.0), f32(1.0))
        vec = [np.full(self.n, init[k], dtype=f32) for k in range(9)]
        cell = {0: (x - 2, y - 2), 1: (3 * x // 4, y - 2), 2: (x // 3, y - 2), 3: (x - 2, y // 2),
                4: (3 * x // 4, y // 2), 5: (x // 2, y // 2), 6: (x - 2, 1), 7: (3 * x // 4, 1),
                8: (x // 2, 1)}.get(int(index))
        if cell is not None:
            vec[int(index)][cell[0] + cell[1] * x] = f32(4.0)
        for b in range(2):
            for k in range(9):
                self.data[b][k][:] = vec[k]
        self.compute_step = 0

    # numeric selectors in the reference's enum order (lbm.rs:10-24): Curl, Ux, Uy, Rho, Speed / Inferno, Viridis, Jet
    STATS = ("curl", "ux", "uy", "rho", "speed")
    CMAPS = ("inferno", "viridis", "jet")

    def compute_summary(self, stat):
        self.summary_stat = self.STATS[int(stat)]
        self.calculate_summary()

    def colors_of(self, cmap):
        self.color_map(self.CMAPS[int(cmap)])
        return self.colors.reshape(self.y, self.x, 3).copy()

    def state(self):
        """everything observable, as arrays shaped (y, x): 18 populations (rest twice = buffer 0), moments,
        output, barrier"""
        shp = (self.y, self.x)
        st = {f"f{b}_{k}": self.population(b, k).reshape(shp).copy() for b in range(2) for k in range(9)}
        st.update(mx=self.ux.reshape(shp).copy(), my=self.uy.reshape(shp).copy(), rho=self.rho.reshape(shp).copy(),
                  out=self.output.reshape(shp).copy(), barrier=self.barrier.reshape(shp).copy())
        return st

    # accessors shaped like oracle/lbm_oracle.py
    def population(self, buffer, k):
        return self.data[0][REST] if k == REST else self.data[buffer][k]


class _Vec3Array(StorageArray):
    """array<vec3<f32>> viewed as n x 3 floats"""

    def load(self, i):
        i = int(i)
        if 0 <= i < len(self.data):
            return self.data[i].copy()
        return np.zeros(3, f32)

    def store(self, i, v):
        i = int(i)
        if 0 <= i < len(self.data):
            self.data[i] = v


def _dispatch_objects(self, workgroups, bindings):
    """like dispatch(), but bindings may already be StorageArray objects (vec3 arrays)"""
    env_globals = {}
    for name, g in self.globals.items():
        b = bindings[(g["group"], g["binding"])]
        if isinstance(b, StorageArray):
            env_globals[name] = b
        elif isinstance(b, np.ndarray):
            env_globals[name] = StorageArray(b)
        else:
            env_globals[name] = b
    with np.errstate(all="ignore"):
        for gid in range(workgroups * self.workgroup_size):
            self._invoke(env_globals, gid)


Shader.dispatch_objects = _dispatch_objects
