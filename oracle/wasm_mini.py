"""TEST INFRASTRUCTURE — a small WebAssembly (MVP + sign-extension + saturating truncation + bulk memory copy/fill)
interpreter, just enough to EXECUTE HOST-SIDE FUNCTIONS OF THE REFERENCE'S OWN SHIPPED BINARY
(/root/reference/lbm-wgpu/pkg/lbm_wgpu_bg.wasm, the wasm-pack build of lbm-wgpu): `set_equil` (lbm.rs:611-643) and
the barrier rasteriser `Line::new` / `Line::new_erased` (barrier_shapes/line.rs:22-87, which carries the un-vendored
`line_drawing 1.0.0` Bresenham).  The reference cannot be built here (no cargo/rustc) and its WebGPU half cannot run
(no browser), but these functions are plain compiled Rust with no imports on their path, so running them gives
outputs of the reference itself for the host-side rows (initial populations, barrier shapes).

tests/golden/make_wasm_golden.py locates the functions in the (name-stripped) binary by their constants and string
references, runs them here and commits the results; the product never imports this module.

Numeric model: i32 / i64 as Python ints (unsigned residues), f32 / f64 as Python floats; f32 results of + - * /
sqrt are computed in double and rounded once to binary32, which is exact for those operations (the double has more
than 2*24+2 significand bits).
"""
import math
import struct


class Trap(Exception):
    pass


def _leb_u(b, p):
    r = s = 0
    while True:
        x = b[p]
        p += 1
        r |= (x & 0x7F) << s
        s += 7
        if not x & 0x80:
            return r, p


def _leb_s(b, p):
    r = s = 0
    while True:
        x = b[p]
        p += 1
        r |= (x & 0x7F) << s
        s += 7
        if not x & 0x80:
            if x & 0x40:
                r -= 1 << s
            return r, p


M32, M64 = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF


def _s32(v):
    return v - (1 << 32) if v & 0x80000000 else v


def _s64(v):
    return v - (1 << 64) if v & (1 << 63) else v


def _f32(x):
    try:
        return struct.unpack("<f", struct.pack("<f", x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


def _trunc(x, lo, hi, sat):
    if x != x:
        if sat:
            return 0
        raise Trap("invalid conversion to integer")
    if math.isinf(x) or not (lo <= math.trunc(x) <= hi):
        if sat:
            return lo if x < 0 else hi
        raise Trap("integer overflow")
    return math.trunc(x)


class Module:
    def __init__(self, data):
        self.bytes = data
        assert data[:8] == b"\x00asm\x01\x00\x00\x00"
        p = 8
        self.types, self.imports, self.func_types, self.exports = [], [], [], {}
        self.globals, self.table, self.bodies, self.segments = [], [], [], []
        self.mem_pages = 0
        while p < len(data):
            sid = data[p]
            size, p = _leb_u(data, p + 1)
            end = p + size
            if sid == 1:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    assert data[p] == 0x60
                    k, p = _leb_u(data, p + 1)
                    params = list(data[p:p + k])
                    p += k
                    k, p = _leb_u(data, p)
                    res = list(data[p:p + k])
                    p += k
                    self.types.append((params, res))
            elif sid == 2:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    k, p = _leb_u(data, p)
                    mod = data[p:p + k].decode()
                    p += k
                    k, p = _leb_u(data, p)
                    name = data[p:p + k].decode()
                    p += k
                    kind = data[p]
                    assert kind == 0, "only function imports are supported"
                    t, p = _leb_u(data, p + 1)
                    self.imports.append((mod, name, t))
            elif sid == 3:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    t, p = _leb_u(data, p)
                    self.func_types.append(t)
            elif sid == 4:
                n, p = _leb_u(data, p)
                assert n == 1 and data[p] == 0x70
                flag = data[p + 1]
                lo, p = _leb_u(data, p + 2)
                if flag & 1:
                    _, p = _leb_u(data, p)
                self.table = [None] * lo
            elif sid == 5:
                n, p = _leb_u(data, p)
                flag = data[p]
                self.mem_pages, p = _leb_u(data, p + 1)
                if flag & 1:
                    _, p = _leb_u(data, p)
            elif sid == 6:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    vt, mut = data[p], data[p + 1]
                    p += 2
                    op = data[p]
                    if op == 0x41:
                        v, p = _leb_s(data, p + 1)
                        v &= M32
                    elif op == 0x42:
                        v, p = _leb_s(data, p + 1)
                        v &= M64
                    else:
                        raise NotImplementedError("global initialiser")
                    assert data[p] == 0x0B
                    p += 1
                    self.globals.append(v)
            elif sid == 7:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    k, p = _leb_u(data, p)
                    name = data[p:p + k].decode()
                    p += k
                    kind = data[p]
                    idx, p = _leb_u(data, p + 1)
                    self.exports[name] = (kind, idx)
            elif sid == 9:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    flag, p = _leb_u(data, p)
                    assert flag == 0 and data[p] == 0x41
                    off, p = _leb_s(data, p + 1)
                    assert data[p] == 0x0B
                    k, p = _leb_u(data, p + 1)
                    for j in range(k):
                        f, p = _leb_u(data, p)
                        self.table[off + j] = f
            elif sid == 10:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    k, p = _leb_u(data, p)
                    self.bodies.append((p, p + k))
                    p += k
            elif sid == 11:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    flag, p = _leb_u(data, p)
                    assert flag == 0 and data[p] == 0x41
                    off, p = _leb_s(data, p + 1)
                    assert data[p] == 0x0B
                    k, p = _leb_u(data, p + 1)
                    self.segments.append((off, data[p:p + k]))
                    p += k
            p = end
        self.n_imports = len(self.imports)
        self._decoded = {}

    def type_of(self, fidx):
        if fidx < self.n_imports:
            return self.types[self.imports[fidx][2]]
        return self.types[self.func_types[fidx - self.n_imports]]

    def body_bytes(self, fidx):
        a, b = self.bodies[fidx - self.n_imports]
        return self.bytes[a:b]

    # -- decode one function body into (nlocals_extra, code) with branch targets resolved --------------------------
    def decode(self, fidx):
        if fidx in self._decoded:
            return self._decoded[fidx]
        b = self.bytes
        p, end = self.bodies[fidx - self.n_imports]
        ngroups, p = _leb_u(b, p)
        local_types = []
        for _ in range(ngroups):
            cnt, p = _leb_u(b, p)
            local_types += [b[p]] * cnt
            p += 1
        code = []
        stack = []  # indices of open block/loop/if instructions
        while p < end:
            op = b[p]
            p += 1
            if op in (0x02, 0x03, 0x04):
                bt = b[p]
                if bt == 0x40:
                    arity, p = 0, p + 1
                elif bt in (0x7F, 0x7E, 0x7D, 0x7C):
                    arity, p = 1, p + 1
                else:
                    t, p = _leb_s(b, p)
                    params, res = self.types[t]
                    assert not params, "multi-value block parameters are not supported"
                    arity = len(res)
                code.append([op, arity, None, None])  # [op, arity, end_pc, else_pc]
                stack.append(len(code) - 1)
            elif op == 0x05:
                code[stack[-1]][3] = len(code)
                code.append([op, None])  # patched with end_pc
            elif op == 0x0B:
                if stack:
                    s = stack.pop()
                    code[s][2] = len(code)
                    if code[s][3] is not None:
                        code[code[s][3]][1] = len(code)
                code.append([op])
            elif op in (0x0C, 0x0D):
                d, p = _leb_u(b, p)
                code.append([op, d])
            elif op == 0x0E:
                n, p = _leb_u(b, p)
                tgt = []
                for _ in range(n + 1):
                    d, p = _leb_u(b, p)
                    tgt.append(d)
                code.append([op, tgt])
            elif op == 0x10:
                f, p = _leb_u(b, p)
                code.append([op, f])
            elif op == 0x11:
                t, p = _leb_u(b, p)
                p += 1
                code.append([op, t])
            elif op in (0x20, 0x21, 0x22, 0x23, 0x24):
                i, p = _leb_u(b, p)
                code.append([op, i])
            elif 0x28 <= op <= 0x3E:
                _, p = _leb_u(b, p)
                off, p = _leb_u(b, p)
                code.append([op, off])
            elif op in (0x3F, 0x40):
                p += 1
                code.append([op])
            elif op == 0x41:
                v, p = _leb_s(b, p)
                code.append([op, v & M32])
            elif op == 0x42:
                v, p = _leb_s(b, p)
                code.append([op, v & M64])
            elif op == 0x43:
                code.append([op, struct.unpack_from("<f", b, p)[0]])
                p += 4
            elif op == 0x44:
                code.append([op, struct.unpack_from("<d", b, p)[0]])
                p += 8
            elif op == 0xFC:
                sub, p = _leb_u(b, p)
                if sub in (10,):
                    p += 2
                elif sub in (11,):
                    p += 1
                elif sub > 7:
                    raise NotImplementedError(f"0xfc {sub}")
                code.append([op, sub])
            else:
                code.append([op])
        self._decoded[fidx] = (local_types, code)
        return self._decoded[fidx]


class Instance:
    def __init__(self, module, imports=None, max_steps=200_000_000, hooks=None):
        self.m = module
        self.mem = bytearray(module.mem_pages * 65536)
        for off, blob in module.segments:
            self.mem[off:off + len(blob)] = blob
        self.globals = list(module.globals)
        self.imports = imports or {}
        self.hooks = hooks or {}  # function index -> callable(instance, *args): replaces a function of the module
        self.virtual_table = {}   # table index beyond the module's table -> callable(instance, *args) for call_indirect
        self.steps = 0
        self.max_steps = max_steps
        self.called = []
        self.stack_trace = []

    # -- memory helpers for the host side
    def read(self, addr, n):
        return bytes(self.mem[addr:addr + n])

    def u32(self, addr):
        return struct.unpack_from("<I", self.mem, addr)[0]

    def i32(self, addr):
        return struct.unpack_from("<i", self.mem, addr)[0]

    def write_u32(self, addr, v):
        struct.pack_into("<I", self.mem, addr, v & M32)

    def call(self, fidx, *args):
        params, res = self.m.type_of(fidx)
        assert len(args) == len(params), (len(args), params)
        out = self._invoke(fidx, list(args))
        return out[0] if len(out) == 1 else (tuple(out) if out else None)

    def run_fragment(self, fidx, start_pc, locals_set):
        """Execute function `fidx` from instruction index `start_pc` with the given locals ({index: value}, the rest
        zero) until control leaves the enclosing blocks (a branch past the fragment's own labels, a return, or the end
        of the body).  For code that rustc inlined into a large function — e.g. a `match` arm of an event loop — and
        that cannot be called on its own: the caller supplies the locals the fragment reads."""
        nparams = len(self.m.type_of(fidx)[0])
        return self._invoke(fidx, [0] * nparams, start_pc=start_pc, locals_set=locals_set)

    def _invoke(self, fidx, args, start_pc=0, locals_set=None):
        m = self.m
        if fidx in self.hooks and not start_pc:
            r = self.hooks[fidx](self, *args)
            return [] if r is None else [r]
        if fidx < m.n_imports:
            mod, name, _ = m.imports[fidx]
            self.called.append(name)
            fn = self.imports.get(name)
            if fn is None:
                raise Trap(f"import {mod}.{name} called (not provided)")
            r = fn(self, *args)
            return [] if r is None else [r]
        params, results = m.type_of(fidx)
        local_types, code = m.decode(fidx)
        self.stack_trace.append(fidx)
        try:
            return self._run(fidx, args, start_pc, locals_set, results, local_types, code)
        except Trap as e:
            if not hasattr(e, "stack"):
                e.stack = list(self.stack_trace)  # function indices, outermost first: where the trap happened
            raise
        finally:
            self.stack_trace.pop()

    def _run(self, fidx, args, start_pc, locals_set, results, local_types, code):
        m = self.m
        loc = args + [0.0 if t in (0x7D, 0x7C) else 0 for t in local_types]
        for k, v in (locals_set or {}).items():
            loc[k] = v
        st = []
        ctl = []  # (is_loop, continuation pc, stack height, arity)
        mem = self.mem
        pc, n = start_pc, len(code)
        pk, up = struct.pack_into, struct.unpack_from
        while pc < n:
            ins = code[pc]
            op = ins[0]
            pc += 1
            self.steps += 1
            if op == 0x20:
                st.append(loc[ins[1]])
            elif op == 0x21:
                loc[ins[1]] = st.pop()
            elif op == 0x22:
                loc[ins[1]] = st[-1]
            elif op == 0x41 or op == 0x42 or op == 0x43 or op == 0x44:
                st.append(ins[1])
            elif op == 0x6A:
                b = st.pop()
                st[-1] = (st[-1] + b) & M32
            elif op == 0x28:
                a = st.pop() + ins[1]
                if a + 4 > len(mem):
                    raise Trap("out of bounds memory access")
                st.append(up("<I", mem, a)[0])
            elif op == 0x36:
                v = st.pop()
                a = st.pop() + ins[1]
                if a + 4 > len(mem):
                    raise Trap("out of bounds memory access")
                pk("<I", mem, a, v)
            elif op == 0x02:
                ctl.append((False, ins[2] + 1, len(st), ins[1]))
            elif op == 0x03:
                ctl.append((True, pc - 1, len(st), 0))
            elif op == 0x04:
                c = st.pop()
                ctl.append((False, ins[2] + 1, len(st), ins[1]))
                if not c:
                    if ins[3] is not None:
                        pc = ins[3] + 1
                    else:
                        pc = ins[2]  # the `end` pops the label
            elif op == 0x05:
                pc = ins[1]  # end of the then-branch: jump to the matching `end`
            elif op == 0x0B:
                if ctl:
                    ctl.pop()
            elif op == 0x0C or op == 0x0D or op == 0x0E:
                if op == 0x0D:
                    if not st.pop():
                        continue
                    d = ins[1]
                elif op == 0x0E:
                    i = st.pop()
                    t = ins[1]
                    d = t[i] if i < len(t) - 1 else t[-1]
                else:
                    d = ins[1]
                if d >= len(ctl):  # branch to the function label = return
                    break
                for _ in range(d):
                    ctl.pop()
                is_loop, cont, height, arity = ctl[-1]
                if arity:
                    vals = st[-arity:]
                    del st[height:]
                    st += vals
                else:
                    del st[height:]
                if is_loop:
                    pc = cont + 1  # re-enter the loop body; the label stays
                else:
                    ctl.pop()
                    pc = cont
            elif op == 0x0F:
                break
            elif op == 0x10:
                f = ins[1]
                np_ = len(m.type_of(f)[0])
                a = st[len(st) - np_:] if np_ else []
                if np_:
                    del st[len(st) - np_:]
                st += self._invoke(f, a)
            elif op == 0x11:
                i = st.pop()
                if i in self.virtual_table:  # a host-provided trait-object method (e.g. a fabricated Rust vtable)
                    np_ = len(m.types[ins[1]][0])
                    a = st[len(st) - np_:] if np_ else []
                    if np_:
                        del st[len(st) - np_:]
                    r = self.virtual_table[i](self, *a)
                    if m.types[ins[1]][1]:
                        st.append(r)
                    continue
                if i >= len(m.table) or m.table[i] is None:
                    raise Trap(f"undefined table element {i:#x} (call_indirect at pc {pc - 1} of function {fidx})")
                f = m.table[i]
                if m.type_of(f) != m.types[ins[1]]:
                    raise Trap("indirect call type mismatch")
                np_ = len(m.types[ins[1]][0])
                a = st[len(st) - np_:] if np_ else []
                if np_:
                    del st[len(st) - np_:]
                st += self._invoke(f, a)
            elif op == 0x1A:
                st.pop()
            elif op == 0x1B:
                c = st.pop()
                b = st.pop()
                if not c:
                    st[-1] = b
            elif op == 0x23:
                st.append(self.globals[ins[1]])
            elif op == 0x24:
                self.globals[ins[1]] = st.pop()
            elif 0x29 <= op <= 0x35:
                a = st.pop() + ins[1]
                fmt, size = _LOADS[op]
                if a + size > len(mem):
                    raise Trap("out of bounds memory access")
                v = up(fmt, mem, a)[0]
                if op in (0x2C, 0x2E):
                    v &= M32
                elif op in (0x30, 0x32, 0x34):
                    v &= M64
                st.append(v)
            elif 0x37 <= op <= 0x3E:
                v = st.pop()
                a = st.pop() + ins[1]
                fmt, size, mask = _STORES[op]
                if a + size > len(mem):
                    raise Trap("out of bounds memory access")
                pk(fmt, mem, a, v & mask if mask else v)
            elif op == 0x3F:
                st.append(len(mem) // 65536)
            elif op == 0x40:
                d = st.pop()
                old = len(mem) // 65536
                if old + d > 16384:
                    st.append(M32)
                else:
                    mem.extend(bytes(d * 65536))
                    st.append(old)
            elif op == 0x00:
                raise Trap("unreachable executed")
            elif op == 0x01:
                pass
            else:
                self._numeric(op, ins, st)
            if self.steps > self.max_steps:
                raise Trap("step budget exhausted")
        nres = len(results)
        return st[len(st) - nres:] if nres else []

    def _numeric(self, op, ins, st):
        if op == 0x45:
            st[-1] = 0 if st[-1] else 1
        elif 0x46 <= op <= 0x4F:
            b, a = st.pop(), st.pop()
            k = op - 0x46
            if k >= 2 and k % 2 == 0:
                a, b = _s32(a), _s32(b)
            st.append(int((a == b, a != b, a < b, a < b, a > b, a > b, a <= b, a <= b, a >= b, a >= b)[k]))
        elif op == 0x50:
            st[-1] = 0 if st[-1] else 1
        elif 0x51 <= op <= 0x5A:
            b, a = st.pop(), st.pop()
            k = op - 0x51
            if k >= 2 and k % 2 == 0:
                a, b = _s64(a), _s64(b)
            st.append(int((a == b, a != b, a < b, a < b, a > b, a > b, a <= b, a <= b, a >= b, a >= b)[k]))
        elif 0x5B <= op <= 0x66:
            b, a = st.pop(), st.pop()
            k = (op - 0x5B) % 6
            st.append(int((a == b, a != b, a < b, a > b, a <= b, a >= b)[k]))
        elif 0x67 <= op <= 0x78:
            self._int_op(op - 0x67, 32, M32, _s32, st)
        elif 0x79 <= op <= 0x8A:
            self._int_op(op - 0x79, 64, M64, _s64, st)
        elif 0x8B <= op <= 0x98:
            self._float_op(op - 0x8B, _f32, st)
        elif 0x99 <= op <= 0xA6:
            self._float_op(op - 0x99, float, st)
        elif op == 0xA7:
            st[-1] &= M32
        elif op in (0xA8, 0xAA):
            st[-1] = _trunc(st[-1], -(1 << 31), (1 << 31) - 1, False) & M32
        elif op in (0xA9, 0xAB):
            st[-1] = _trunc(st[-1], 0, M32, False)
        elif op == 0xAC:
            st[-1] = _s32(st[-1]) & M64
        elif op == 0xAD:
            pass
        elif op in (0xAE, 0xB0):
            st[-1] = _trunc(st[-1], -(1 << 63), (1 << 63) - 1, False) & M64
        elif op in (0xAF, 0xB1):
            st[-1] = _trunc(st[-1], 0, M64, False)
        elif op == 0xB2:
            st[-1] = _f32(float(_s32(st[-1])))
        elif op == 0xB3:
            st[-1] = _f32(float(st[-1]))
        elif op == 0xB4:
            st[-1] = _f32_from_int(_s64(st[-1]))
        elif op == 0xB5:
            st[-1] = _f32_from_int(st[-1])
        elif op == 0xB6:
            st[-1] = _f32(st[-1])
        elif op == 0xB7:
            st[-1] = float(_s32(st[-1]))
        elif op == 0xB8:
            st[-1] = float(st[-1])
        elif op == 0xB9:
            st[-1] = float(_s64(st[-1]))
        elif op == 0xBA:
            st[-1] = float(st[-1])
        elif op == 0xBB:
            pass
        elif op == 0xBC:
            st[-1] = struct.unpack("<I", struct.pack("<f", st[-1]))[0]
        elif op == 0xBD:
            st[-1] = struct.unpack("<Q", struct.pack("<d", st[-1]))[0]
        elif op == 0xBE:
            st[-1] = struct.unpack("<f", struct.pack("<I", st[-1]))[0]
        elif op == 0xBF:
            st[-1] = struct.unpack("<d", struct.pack("<Q", st[-1]))[0]
        elif op == 0xC0:
            v = st[-1] & 0xFF
            st[-1] = (v - 256 if v & 0x80 else v) & M32
        elif op == 0xC1:
            v = st[-1] & 0xFFFF
            st[-1] = (v - 65536 if v & 0x8000 else v) & M32
        elif op == 0xC2:
            v = st[-1] & 0xFF
            st[-1] = (v - 256 if v & 0x80 else v) & M64
        elif op == 0xC3:
            v = st[-1] & 0xFFFF
            st[-1] = (v - 65536 if v & 0x8000 else v) & M64
        elif op == 0xC4:
            st[-1] = _s32(st[-1] & M32) & M64
        elif op == 0xFC:
            sub = ins[1]
            if sub <= 7:
                bits = 32 if sub < 4 else 64
                signed = sub % 2 == 0
                lo, hi = (-(1 << (bits - 1)), (1 << (bits - 1)) - 1) if signed else (0, (1 << bits) - 1)
                st[-1] = _trunc(st[-1], lo, hi, True) & ((1 << bits) - 1)
            elif sub == 10:
                n, s, d = st.pop(), st.pop(), st.pop()
                if s + n > len(self.mem) or d + n > len(self.mem):
                    raise Trap("out of bounds memory access")
                self.mem[d:d + n] = self.mem[s:s + n]
            elif sub == 11:
                n, v, d = st.pop(), st.pop(), st.pop()
                if d + n > len(self.mem):
                    raise Trap("out of bounds memory access")
                self.mem[d:d + n] = bytes([v & 0xFF]) * n
        else:
            raise NotImplementedError(f"opcode 0x{op:02x}")

    @staticmethod
    def _int_op(k, bits, mask, signed, st):
        if k == 0:  # clz
            v = st[-1]
            st[-1] = bits - v.bit_length()
            return
        if k == 1:  # ctz
            v = st[-1]
            st[-1] = bits if v == 0 else (v & -v).bit_length() - 1
            return
        if k == 2:
            st[-1] = bin(st[-1]).count("1")
            return
        b, a = st.pop(), st.pop()
        if k == 3:
            r = a + b
        elif k == 4:
            r = a - b
        elif k == 5:
            r = a * b
        elif k == 6:  # div_s
            if b == 0:
                raise Trap("integer divide by zero")
            sa, sb = signed(a), signed(b)
            if sa == -(1 << (bits - 1)) and sb == -1:
                raise Trap("integer overflow")
            r = abs(sa) // abs(sb)
            if (sa < 0) != (sb < 0):
                r = -r
        elif k == 7:
            if b == 0:
                raise Trap("integer divide by zero")
            r = a // b
        elif k == 8:  # rem_s
            if b == 0:
                raise Trap("integer divide by zero")
            sa, sb = signed(a), signed(b)
            r = abs(sa) % abs(sb)
            if sa < 0:
                r = -r
        elif k == 9:
            if b == 0:
                raise Trap("integer divide by zero")
            r = a % b
        elif k == 10:
            r = a & b
        elif k == 11:
            r = a | b
        elif k == 12:
            r = a ^ b
        elif k == 13:
            r = a << (b % bits)
        elif k == 14:
            r = signed(a) >> (b % bits)
        elif k == 15:
            r = a >> (b % bits)
        elif k == 16:
            s = b % bits
            r = (a << s) | (a >> (bits - s)) if s else a
        elif k == 17:
            s = b % bits
            r = (a >> s) | (a << (bits - s)) if s else a
        else:
            raise NotImplementedError
        st.append(r & mask)

    @staticmethod
    def _float_op(k, rnd, st):
        if k <= 6:
            a = st[-1]
            if k == 0:
                r = abs(a)
            elif k == 1:
                r = -a
            elif k == 2:
                r = float(math.ceil(a)) if math.isfinite(a) else a
            elif k == 3:
                r = float(math.floor(a)) if math.isfinite(a) else a
            elif k == 4:
                r = float(math.trunc(a)) if math.isfinite(a) else a
            elif k == 5:
                r = float(round(a)) if math.isfinite(a) else a  # Python rounds half to even, as wasm `nearest`
            else:
                r = math.sqrt(a) if a >= 0 else math.nan
            if r == 0 and k in (2, 3, 4, 5):
                r = math.copysign(0.0, a)
            st[-1] = rnd(r)
            return
        b, a = st.pop(), st.pop()
        if k == 7:
            r = a + b
        elif k == 8:
            r = a - b
        elif k == 9:
            r = a * b
        elif k == 10:
            if b == 0:
                r = math.nan if (a == 0 or a != a) else math.copysign(math.inf, a) * math.copysign(1.0, b)
            else:
                r = a / b
        elif k == 11:
            r = math.nan if (a != a or b != b) else (min(a, b) if a != b else (a if math.copysign(1, a) < 0 else b))
        elif k == 12:
            r = math.nan if (a != a or b != b) else (max(a, b) if a != b else (a if math.copysign(1, a) > 0 else b))
        else:
            r = math.copysign(a, b)
        st.append(rnd(r))


def _f32_from_int(v):
    # int -> f32 with a single rounding (float(v) would round to double first)
    import numpy as np
    return float(np.float32(v)) if abs(v) < (1 << 53) else float(np.array(v, dtype=np.longdouble).astype(np.float32))


_LOADS = {0x29: ("<Q", 8), 0x2A: ("<f", 4), 0x2B: ("<d", 8), 0x2C: ("<b", 1), 0x2D: ("<B", 1), 0x2E: ("<h", 2),
          0x2F: ("<H", 2), 0x30: ("<b", 1), 0x31: ("<B", 1), 0x32: ("<h", 2), 0x33: ("<H", 2), 0x34: ("<i", 4),
          0x35: ("<I", 4)}
_STORES = {0x37: ("<Q", 8, 0), 0x38: ("<f", 4, 0), 0x39: ("<d", 8, 0), 0x3A: ("<B", 1, 0xFF), 0x3B: ("<H", 2, 0xFFFF),
           0x3C: ("<B", 1, 0xFF), 0x3D: ("<H", 2, 0xFFFF), 0x3E: ("<I", 4, M32)}
