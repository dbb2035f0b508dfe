"""The C++ cone plugin concept as a public boundary (reference conex/constraint.h:51-197,
conex/cone_program.h:191-218): a cone type the library has never seen, written against include/conex_b200/*.h
only (tests/plugin/out_of_tree_cone.cc), is compiled by plain g++, linked with lib/libconex_b200_host.a and solved
through Program::AddConstraint. CPU: it compiles, links, and fails loudly without a device (no CPU fallback).
GPU: it reaches the same solution as the library's own LinearConstraint and as the oracle."""
import os
import subprocess

import numpy as np
import pytest

from harness import ROOT, oracle

SRC = os.path.join(ROOT, "tests", "plugin", "out_of_tree_cone.cc")
EXE = os.path.join(ROOT, "tests", "plugin", "_build", "out_of_tree_cone")
ARCHIVE = os.path.join(ROOT, "conex_b200", "lib", "libconex_b200_host.a")


def build():
    assert os.path.exists(ARCHIVE), f"{ARCHIVE} missing: run __graft_entry__.build()"
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(ARCHIVE)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I",
                               "/usr/local/cuda/include", SRC, ARCHIVE, "-L/usr/local/cuda/lib64", "-lcudart_static",
                               "-ldl", "-lrt", "-lpthread", "-o", EXE])
    return EXE


def test_out_of_tree_cone_compiles_against_public_headers_only():
    exe = build()
    # the translation unit includes nothing from conex_b200/csrc
    text = open(SRC).read()
    assert "csrc" not in text.split("#include", 1)[1].split("namespace user")[0]
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode != 0 and "conex-b200" in (r.stderr + r.stdout)   # loud failure, no CPU path


@pytest.mark.gpu
def test_out_of_tree_cone_solves_like_the_library_cone_and_the_oracle():
    exe = build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows = {}
    for line in r.stdout.strip().splitlines():
        f = line.split()
        rows[f[0]] = dict(its=int(f[2]), y=np.array([float(f[4]), float(f[5])]), x=np.array([float(v) for v in f[7:]]))
    a, b = rows["out_of_tree"], rows["library"]
    assert abs(a["its"] - b["its"]) <= 1
    assert np.abs(a["y"] - b["y"]).max() <= 1e-9 * np.abs(b["y"]).max()
    assert np.abs(a["x"] - b["x"]).max() <= 1e-7 * max(1.0, np.abs(b["x"]).max())
    # the same LP on the oracle
    A = np.array([[1, 3], [4, 1], [1, 1], [0.3, -0.2], [-0.7, 0.5]], dtype=np.float64)
    c = np.array([1, 1, 1, 2, 1.5])
    O = oracle()
    P = O.program()
    P.add_linear(A, c)
    solved, yo = P.maximize([6.0, 5.0], O.default_config(prepare_dual_variables=1))
    assert solved == 1
    assert np.abs(a["y"] - yo).max() <= 1e-7 * np.abs(yo).max()
    assert np.abs(a["x"] - P.dual_variable(0).ravel()).max() <= 1e-6
