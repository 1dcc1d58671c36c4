"""CPU suite, part 2: the C-ABI library loads and exports every symbol that
include/pogs_b200.h declares; host-side logic of the Python surface.  No compute
call is made without a GPU (they must fail loudly instead)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pogs_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b(Pogs[A-Za-z]+|pogs_b200_[a-z_0-9]+)\s*\(", txt)
    return sorted(set(names))


def test_header_declares_reference_entry_points():
    names = declared_symbols()
    for must in ("PogsD", "PogsS", "PogsSparseD", "PogsSparseS"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from pogs_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_reference_named_copy_exists():
    """The same build is installed under the file name the reference's graph.py searches
    for (python/pogs/graph.py:29-67)."""
    assert os.path.exists(os.path.join(ROOT, "pogs_b200", "lib", "libpogs_cpu.so"))


def test_enum_abi_values():
    """Pinned by the reference's tests/test_c_interface.cpp:149-162."""
    from pogs_b200 import Function, Ordering

    assert Function.kAbs == 0 and Function.kSquare == 14 and Function.kZero == 15
    assert Function.kLogistic == 8 and Function.kMaxPos0 == 10 and Function.kIndGe0 == 6
    assert Ordering.COL_MAJ == 0 and Ordering.ROW_MAJ == 1
    txt = open(HEADER).read()
    order = re.search(r"enum FUNCTION \{([^}]*)\}", txt).group(1).replace("\n", " ").split(",")
    order = [o.strip() for o in order]
    assert order.index("ABS") == 0 and order.index("SQUARE") == 14 and order.index("ZERO") == 15


def test_function_vector_from_objects_and_encodings():
    from pogs_b200 import Function, FunctionObj, FunctionVector
    from pogs_b200 import graph

    objs = [FunctionObj(Function.kSquare, 1.0, float(i), 1.0) for i in range(5)]
    fv = FunctionVector.from_any(objs)
    assert fv.size == 5 and fv.h.tolist() == [14] * 5 and fv.b.tolist() == [0, 1, 2, 3, 4]
    a, b, c, d, e, h = fv.arrays(np.float32)
    assert a.dtype == np.float32 and h.dtype == np.int32 and a.flags.c_contiguous
    # canonical encodings (reference graph.py:428-431, 520-522, 561-568, 614-620, 660-663, 702-705)
    bb = np.array([1.0, -1.0, 1.0])
    f, g = graph.lasso_functions(3, 2, bb, 0.5)
    assert f.h.tolist() == [14] * 3 and f.b.tolist() == bb.tolist() and g.h.tolist() == [0, 0] and g.c.tolist() == [0.5, 0.5]
    f, g = graph.elastic_net_functions(3, 2, bb, 0.5, 0.2)
    assert g.e.tolist() == [0.1, 0.1]
    f, g = graph.logistic_functions(3, 2, bb, 0.0)
    assert f.h.tolist() == [8] * 3 and f.a.tolist() == (-bb).tolist() and g.h.tolist() == [15, 15]
    f, g = graph.huber_functions(3, 2, bb, 2.0, 0.1)
    assert f.h.tolist() == [2] * 3 and f.a.tolist() == [0.5] * 3 and f.b.tolist() == (bb / 2).tolist() and f.c.tolist() == [4.0] * 3
    f, g = graph.svm_functions(3, 2, bb, 1.5)
    assert f.h.tolist() == [10] * 3 and f.b.tolist() == [-1.0] * 3 and g.h.tolist() == [14, 14] and g.c.tolist() == [1.5, 1.5]
    f, g = graph.nonneg_ls_functions(3, 2, bb)
    assert g.h.tolist() == [6, 6]


def test_fails_loudly_without_gpu():
    """No CPU fallback: without a device the entry points report POGS_ERROR (6) and the
    Python wrapper raises."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    import pogs_b200

    A = np.random.default_rng(0).standard_normal((8, 3))
    with pytest.raises(RuntimeError):
        pogs_b200.solve_lasso(A, np.ones(8), 0.1)
    with pytest.raises(RuntimeError):
        pogs_b200.Solver(A)


def _load_reference_wrapper(tmp_path):
    """The reference's own python/pogs/graph.py, unmodified, pointed at THIS library: the module is
    symlinked (not copied) into a scratch package directory next to a link to pogs_b200/lib/libpogs_cpu.so,
    which is one of the places its _find_shared_library() looks (graph.py:29-67)."""
    import importlib.util

    ref = os.environ.get("POGS_REFERENCE_DIR", "/root/reference")
    src = os.path.join(ref, "python", "pogs", "graph.py")
    if not os.path.exists(src):
        pytest.skip("reference checkout not present (set POGS_REFERENCE_DIR)")
    pkg = tmp_path / "pogs_ref_pkg"
    pkg.mkdir()
    os.symlink(src, pkg / "graph.py")
    os.symlink(os.path.join(ROOT, "pogs_b200", "lib", "libpogs_cpu.so"), pkg / "libpogs_cpu.so")
    spec = importlib.util.spec_from_file_location("pogs_ref_graph", str(pkg / "graph.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)   # binds PogsD / PogsSparseD with argtypes: fails if a symbol is missing
    return mod


def test_reference_python_wrapper_binds_against_this_library(tmp_path):
    """Drop-in check of SURVEY 8b: the unmodified reference wrapper imports against libpogs_cpu.so built
    here (PogsD and PogsSparseD are bound unconditionally at import, graph.py:167-233)."""
    mod = _load_reference_wrapper(tmp_path)
    assert os.path.realpath(mod._lib_path) == os.path.realpath(os.path.join(ROOT, "pogs_b200", "lib", "libpogs_cpu.so"))
    assert mod._lib.PogsD.restype is ctypes.c_int and mod._lib.PogsSparseD.restype is ctypes.c_int
    assert int(mod.Function.kSquare) == 14 and int(mod.Ordering.ROW_MAJ) == 1


def test_reference_python_wrapper_solves_through_this_library(tmp_path):
    """With a GPU: the reference's solve_lasso, unmodified, runs on the device path and returns the
    reference's own result (golden vector).  Without one: the call comes back with POGS_ERROR (6) --
    there is no CPU fallback to hide behind."""
    mod = _load_reference_wrapper(tmp_path)
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import problems

    p = problems.build("c1_lasso_500x300")
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    r = mod.solve_lasso(p["A"], p["b"], p["lam"])
    if not have_gpu:
        assert r["status"] == 6
        return
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))
    assert r["status"] == 0
    assert np.linalg.norm(r["x"] - g["c1_lasso_500x300/float64/x"]) <= 5e-4 * np.linalg.norm(g["c1_lasso_500x300/float64/x"])


def test_cone_enum_abi_values():
    """Pinned by the reference's tests/test_c_interface.cpp:157-162."""
    from pogs_b200 import Cone

    assert Cone.ZERO == 0 and Cone.NON_NEG == 1 and Cone.NON_POS == 2 and Cone.SOC == 3
    txt = open(HEADER).read()
    order = [o.strip() for o in re.search(r"enum CONE \{([^}]*)\}", txt).group(1).split(",")]
    assert order[:4] == ["CONE_ZERO", "CONE_NON_NEG", "CONE_NON_POS", "CONE_SOC"]


def test_reference_cone_wrapper_binds_against_this_library(tmp_path):
    """The unmodified reference python/pogs_cone.py binds PogsConeD / QD / DirectD / DirectQD at import
    (pogs_cone.py:76-180) and looks for the library under <root>/build/lib (pogs_cone.py:18-31)."""
    import importlib.util

    ref = os.environ.get("POGS_REFERENCE_DIR", "/root/reference")
    src = os.path.join(ref, "python", "pogs_cone.py")
    if not os.path.exists(src):
        pytest.skip("reference checkout not present (set POGS_REFERENCE_DIR)")
    (tmp_path / "python").mkdir(); (tmp_path / "build" / "lib").mkdir(parents=True)
    os.symlink(src, tmp_path / "python" / "pogs_cone.py")
    os.symlink(os.path.join(ROOT, "pogs_b200", "lib", "libpogs_cpu.so"), tmp_path / "build" / "lib" / "libpogs_cpu.so")
    spec = importlib.util.spec_from_file_location("pogs_ref_cone", str(tmp_path / "python" / "pogs_cone.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert int(mod.Cone.NON_NEG) == 1
    r = mod.solve_cone(np.array([[1.0, 1.0]]), np.array([2.0]), np.array([1.0, 0.0]), [(mod.Cone.NON_NEG, [0, 1])],
                       [(mod.Cone.ZERO, [0])], max_iter=1000)
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        assert r["status"] == 0 and abs(r["x"][1] - 2.0) < 0.01
    else:
        assert r["status"] == 6   # no CPU fallback


def _plan(rows, cols, nnz, sms=148, elem=4):
    from pogs_b200 import _lib

    out = (ctypes.c_ulonglong * 8)()
    rc = _lib.lib.pogs_b200_plan_sparse_tiles(rows, cols, nnz, sms, elem, out)
    keys = ("P", "Q", "tr", "tc", "ns", "tiles", "smem", "limit")
    return rc, dict(zip(keys, [int(v) for v in out]))


@pytest.mark.parametrize("rows,cols,nnz,elem", [
    (1000000, 100000, 100000000, 4),    # C5, the CSR copy
    (100000, 1000000, 100000000, 4),    # C5, the copy of the transpose
    (1000000, 100000, 100000000, 8),
    (100000, 10000, 1000000, 4),        # the scaled-down twin
    (2000, 300, 12000, 8),
    (700, 260, 5427, 4),
    (120000, 50000, 960000, 4),
    (50000, 120000, 960000, 8),
    (5, 3, 7, 4),
    (3, 4000000, 12, 4),
])
def test_sparse_tile_planner_invariants(rows, cols, nnz, elem):
    """Host logic of the 2-D tiled sparse layout (sparse_tiled.cuh: plan_tiled): the tiles cover the matrix, the
    slice of the multiplied vector of one tile fits in shared memory next to the TMA ring, local row / column
    indices fit 16 bits, no column tile is empty."""
    rc, p = _plan(rows, cols, nnz, 148, elem)
    assert rc == 0
    assert p["P"] * p["tr"] >= rows and (p["P"] - 1) * p["tr"] < rows
    assert p["Q"] * p["tc"] >= cols and (p["Q"] - 1) * p["tc"] < cols
    assert p["tiles"] == p["P"] * p["Q"]
    assert p["tc"] % 32 == 0 and p["tc"] <= 65536 and p["tr"] <= 65504
    assert p["ns"] == (p["tr"] + 31) // 32
    assert p["smem"] <= p["limit"] <= 227 * 1024


def test_sparse_tile_planner_fills_the_gpu_on_the_baseline_shape():
    """C5 (1M x 100k, 1e8 entries, fp32) on 148 SMs: one tile per SM for both copies, the column tiles as few as
    the shared-memory budget allows (their count is the number of partial sums per row)."""
    rc, a = _plan(1000000, 100000, 100000000)
    rc2, t = _plan(100000, 1000000, 100000000)
    assert rc == 0 and rc2 == 0
    assert a["tiles"] == 148 and t["tiles"] == 148
    assert (a["P"], a["Q"]) == (37, 4) and (t["P"], t["Q"]) == (4, 37)


def test_sparse_tile_planner_rejects_what_it_cannot_tile():
    rc, _ = _plan(0, 10, 0)
    assert rc == 1
    rc, _ = _plan(10, 128 * 65536 * 4, 100)   # more column tiles than a row's scatter list holds
    assert rc == 1
