"""The compiled-language host layer: include/blbm.hpp (C++ mirror of lbm-wgpu's `pub struct LBM` and barrier
shapes, standing in for the Rust crate that cannot be built here) drives the library from a C++ program that
compares against the C oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_lbm")


def build():
    src = os.path.join(ROOT, "tests", "cpp", "test_lbm.cpp")
    libdir, oradir = os.path.join(ROOT, "lbm_b200"), os.path.join(ROOT, "oracle")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
           "-L", libdir, "-lblbm", "-L", oradir, "-llbm_oracle", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{oradir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_cpp_wrapper_compiles_links_and_fails_loudly_without_gpu():
    exe = build()
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_wrapper_parity_with_oracle():
    exe = build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bit-identical" in r.stdout
