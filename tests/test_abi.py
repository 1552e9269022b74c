"""The C-ABI library loads and exports every symbol include/blbm.h declares; the ctypes prototype table
of the Python mirror covers exactly the same set; without a GPU every entry point fails loudly (there is
no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re

import pytest

import lbm_b200
from lbm_b200 import lbm as host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "blbm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(blbm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_surface():
    syms = declared_symbols()
    # one entry point per public method of `pub struct LBM` (lbm.rs:726-1515) that is on the hot path
    for name in ("blbm_create", "blbm_destroy", "blbm_iterate", "blbm_collide", "blbm_stream", "blbm_rerender",
                 "blbm_set_summary", "blbm_set_omega", "blbm_reset_to_equilibrium", "blbm_custom_speed",
                 "blbm_single_cell", "blbm_draw_points", "blbm_reset_barrier", "blbm_get_compute_num",
                 "blbm_get_frame_num", "blbm_read_population", "blbm_read_moments", "blbm_read_output",
                 "blbm_read_barrier", "blbm_read_cell_class"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    path = lbm_b200.library_path()
    assert os.path.exists(path), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    lib = C.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in blbm.h but not exported: {missing}"


def test_python_prototypes_cover_the_header_exactly():
    assert sorted(host.PROTOTYPES) == declared_symbols()
    lib = lbm_b200.load_library()
    assert lib.blbm_abi_version() == 1
    assert isinstance(lib.blbm_last_error(), bytes)


def test_library_has_no_link_time_dependency_on_the_driver_or_torch():
    """dlopen must work on a machine without libcuda.so.1 (this container): cudart is linked statically."""
    import subprocess
    out = subprocess.run(["ldd", lbm_b200.library_path()], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out and "torch" not in out


def test_without_a_gpu_the_product_fails_loudly():
    lib = lbm_b200.load_library()
    n = lib.blbm_device_count()
    if n > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(lbm_b200.BlbmError) as e:
        lbm_b200.LBM(1.25, 64, 32)
    assert e.value.code == -4  # BLBM_ENOGPU
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under lbm_b200/ or include/ may reference it."""
    bad = []
    for base in ("lbm_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    if "oracle" in open(os.path.join(dirpath, f), errors="ignore").read().lower():
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_slab_rows_partition():
    from lbm_b200.lbm import slab_rows
    for y, n in ((50, 3), (16384 * 8, 8), (7, 2), (65536, 8)):
        r = slab_rows(y, n)
        assert r[0][0] == 0 and r[-1][1] == y
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_packed_adds_are_never_contracted_into_packed_fmas():
    """The step kernels collide cell pairs with sm_100's packed fp32 adds (FADD2).  ptxas 12.9 contracts
    mul.rn.f32x2 + add.rn.f32x2 into FFMA2 despite the explicit roundings, so the multiplies stay scalar; the
    shipped SASS must hold packed adds and no packed multiply / multiply-add (parity contract: every op
    individually rounded)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lbm_b200.library_path()], capture_output=True, text=True).stdout
    assert "FADD2" in sass
    assert "FFMA2" not in sass and "FMUL2" not in sass


def test_rust_sys_crate_declares_every_entry_point():
    """rust/blbm-sys is shipped as source only (no Rust toolchain in this image): at least keep it complete —
    one `pub fn` per symbol of include/blbm.h — and its tuning constants in step with the header's enum."""
    rs = open(os.path.join(ROOT, "rust", "blbm-sys", "src", "lib.rs")).read()
    missing = [s for s in declared_symbols() if not re.search(r"\bfn\s+" + s + r"\b", rs)]
    assert not missing, missing
    hdr = open(os.path.join(ROOT, "include", "blbm.h")).read()
    for name, val in re.findall(r"\b(BLBM_TUNE_[A-Z0-9_]+)\s*=\s*(\d+)", hdr):
        assert re.search(r"\b" + name + r":\s*c_int\s*=\s*" + val + r"\b", rs), f"{name} = {val} missing in blbm-sys"


def test_default_step_kernels_keep_their_register_budget_and_wide_accesses():
    """Static guard for the measured configuration (no GPU needed): the instantiations the default launch paths
    use — step_vec4_kernel<no moments, 4 rows, flavour 0 | 2, scalar adds, 32-bit offsets> — must stay at
    64 registers without a spill (8 blocks of 128 threads per SM), move the populations with 128-bit global
    accesses, stage the bounce-back's own rows with cp.async (flavour 2), realign the x+-1 gathers with warp
    shuffles.  profiles/r2/sass_summary.txt is the same table for every kernel of the library."""
    import importlib.util
    import shutil
    if not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("cuobjdump not available")
    spec = importlib.util.spec_from_file_location("sass_summary", os.path.join(ROOT, "profiles", "sass_summary.py"))
    ss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ss)
    res = ss.resource_usage()
    counts, _ = ss.sass_counts()
    names = {ss.short(v): k for k, v in ss.demangle(sorted(res)).items()}
    for flavour in (0, 2):
        k = names[f"step_vec4_kernel<(bool)0, (int)4, (int){flavour}, (bool)0, u32, (bool)0>"]
        assert int(res[k]["REG"]) <= 64 and int(res[k]["STACK"]) == 0, (flavour, res[k])
        # linked slabs (handshake inside the kernel): same occupancy; two block-uniform words may live on the stack
        kl = names[f"step_vec4_kernel<(bool)0, (int)4, (int){flavour}, (bool)0, u32, (bool)1>"]
        assert int(res[kl]["REG"]) <= 64 and int(res[kl]["STACK"]) <= 8, (flavour, res[kl])
        c = counts[k]
        assert c["LDG.E.128"] >= 9 and c["STG.E.128"] >= 9 and c["SHFL"] >= 6 and not c["LDL"] and not c["STL"]
        assert (c["LDGSTS"] > 0) == (flavour == 2)
    # the moment-storing launch (one per iterate) stores its moments before the collision: same occupancy, at most
    # four block-uniform / address words on the stack
    for flavour in (0, 2):
        k = names[f"step_vec4_kernel<(bool)1, (int)4, (int){flavour}, (bool)0, u32, (bool)0>"]
        assert int(res[k]["REG"]) <= 64 and int(res[k]["STACK"]) <= 16, (flavour, res[k])


def test_every_entry_point_rejects_a_null_handle_without_crashing():
    """Error behaviour at the boundary: the reference panics (expect/unwrap); the C ABI must return a status.
    Every handle-taking entry point called with NULL (and zero / NULL for the rest) returns a negative blbm_status —
    or the neutral value for the plain getters — sets blbm_last_error, and never dereferences the handle."""
    lib = lbm_b200.load_library()
    getters = {"blbm_get_compute_num": 0, "blbm_get_frame_num": 0, "blbm_get_launch_count": 0,
               "blbm_get_device_bytes": 0, "blbm_get_lazy_barriers_active": 0, "blbm_group_size": 0}
    no_handle = {"blbm_last_error", "blbm_abi_version", "blbm_device_count", "blbm_create", "blbm_create_slab",
                 "blbm_create_group", "blbm_rasterize_line", "blbm_preset_lines"}
    checked = 0
    for name, (res, args) in host.PROTOTYPES.items():
        if name in no_handle:
            continue
        assert args and args[0] is host._P, name
        zeros = [a() if a is not host._P and not hasattr(a, "contents") else None for a in args]
        rc = getattr(lib, name)(*zeros)
        if name == "blbm_destroy":
            assert rc == 0  # destroying nothing is fine, like dropping an Option::None
        elif name in getters:
            assert rc == getters[name], name
        else:
            assert rc < 0, f"{name}(NULL, ...) returned {rc}"
            assert lib.blbm_last_error(), name
        checked += 1
    assert checked >= 45
    # creation with bad arguments: status, no handle
    h = host._P()
    assert lib.blbm_create(0, 8, 1.0, 0.1, 0, C.byref(h)) < 0 and not h
    assert lib.blbm_create_slab(8, 8, 4, 4, 1.0, 0.1, 0, C.byref(h)) < 0 and not h
    assert lib.blbm_create(8, 8, 1.0, 0.1, 0, None) < 0
    devs = (C.c_int * 2)(0, 0)
    assert lib.blbm_create_group(8, 8, 1.0, 0.1, None, 2, C.byref(h)) < 0 and not h
    assert lib.blbm_create_group(8, 8, 1.0, 0.1, devs, 0, C.byref(h)) < 0 and not h
    assert lib.blbm_create_group(8, 3, 1.0, 0.1, devs, 2, C.byref(h)) < 0 and not h  # fewer than 2 rows per slab


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/blbm.h must compile as C99 (what cgo / bindgen / a Rust -sys crate's
    build script would feed it to), and a C program using it must link against the library."""
    import subprocess
    import tempfile
    hdr = os.path.join(ROOT, "include", "blbm.h")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = ('#include "blbm.h"\n#include <stdio.h>\n'
           "int main(void) { blbm_t *h = 0; int rc = blbm_create(64, 32, 1.25f, 0.1f, 0, &h);\n"
           '  printf("%d %d %s\\n", blbm_abi_version(), rc, blbm_last_error()); if (h) blbm_destroy(h); return 0; }\n')
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(c, "w").write(src)
        libdir = os.path.dirname(lbm_b200.library_path())
        r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe, "-L", libdir, "-lblbm",
                            f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and r.stdout.startswith("1 "), r.stdout + r.stderr
