"""The C ABI: libcpab_b200.so loads on a machine without a GPU and exports exactly the symbols
include/libcpab_b200.h declares; argument errors are reported through status codes and
cpab_b200_last_error(), never by exiting.  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "libcpab_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cpab_b200_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from libcpab_b200 import _lib
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    from libcpab_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_every_declaration_cites_the_reference_interface_it_replaces():
    text = open(HEADER).read()
    for fn in ("findcellidx", "theta_to_trels", "forward", "backward_jacobian", "backward_theta",
               "interpolate_forward"):
        block = text[:text.index("int cpab_b200_" + fn + "(")]
        comment = block[block.rindex("/*"):]
        assert re.search(r"libcpab/[\w/]+\.(py|cpp|cu):\d+", comment), fn


def test_version_and_build_info(lib):
    assert lib.cpab_b200_abi_version() == 2
    info = lib.cpab_b200_build_info().decode()
    assert "sm_100a" in info


def test_argument_errors_return_codes_and_messages(lib):
    nc = (ctypes.c_int * 2)(3, 3)
    bad_nc = (ctypes.c_int * 2)(3, 0)
    rc = lib.cpab_b200_forward(0, 0, 5, nc, 50, 1, 1, 0, None, None, None, None)
    assert rc == -1 and b"ndim" in lib.cpab_b200_last_error()
    rc = lib.cpab_b200_forward(0, 0, 2, bad_nc, 50, 1, 1, 0, None, None, None, None)
    assert rc == -1 and b"positive" in lib.cpab_b200_last_error()
    rc = lib.cpab_b200_forward(0, 0, 2, nc, 0, 1, 1, 0, None, None, None, None)
    assert rc == -1 and b"nstepsolver" in lib.cpab_b200_last_error()
    rc = lib.cpab_b200_forward(0, 0, 2, nc, 50, 1, 8, 0, None, None, None, None)
    assert rc == -1 and b"NULL" in lib.cpab_b200_last_error()
    rc = lib.cpab_b200_forward(7, 0, 2, nc, 50, 1, 8, 0, None, None, None, None)
    assert rc == -1 and b"dtype" in lib.cpab_b200_last_error()
    assert lib.cpab_b200_set_tuning(b"bogus", 3) == -1
    # workspace query is pure host arithmetic: G [n_theta, D] + RK2 step records [n_theta, nC, 8] (2-D)
    # + certificate bounds [n_theta, 2] float + 16 bytes of work counters + (float32) 1 bit per trajectory
    assert lib.cpab_b200_backward_workspace_bytes(0, 2, nc, 10, 1000) == 10 * 36 * (6 + 8) * 4 + 10 * 8 + 16 + 10 * 32 * 4
    assert lib.cpab_b200_backward_workspace_bytes(1, 2, nc, 10, 1000) == 10 * 36 * (6 + 8) * 8 + 10 * 8 + 16
    # empty problems are accepted without touching the device
    assert lib.cpab_b200_forward(0, 0, 2, nc, 50, 0, 8, 0, None, None, None, None) == 0
    assert lib.cpab_b200_launch_count() >= 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "libcpab_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libcpab_oracle" not in src and "libcpab_ref" not in src, f
