"""CPU-side checks: the C-ABI library loads and exports every symbol include/bnr.h declares (no compute calls
without a GPU), the ctypes prototypes cover the header, and the host-side Fit!/Summary logic (Summary tables, input
formats) behaves like the reference.  The control loops of bnr_fit are checked in tests/test_fit_plan.py."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import bnr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "bnr.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bnr_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(bnr):
    lib = ctypes.CDLL(os.path.join(ROOT, "bayesiannetworkregression.jl_b200", "libbnr.so"))
    syms = _header_symbols()
    assert len(syms) >= 30
    for name in syms:
        assert hasattr(lib, name), "libbnr.so does not export " + name
    from bnr_b200 import capi
    assert sorted(capi.PROTOTYPES) == syms, "ctypes prototypes and include/bnr.h disagree"
    assert bnr.lib().bnr_version() == 200


def test_params_struct_matches_header(bnr):
    from bnr_b200 import capi
    p = capi.Params()
    bnr.lib().bnr_default_params(ctypes.byref(p))
    assert (p.eta, p.zeta, p.iota, p.a_delta, p.b_delta, p.nu) == (1.01, 1.0, 1.0, 1.0, 1.0, 10.0)
    assert p.num_chains == 2 and p.gig_inject_len == 64 and p.trace_full_chains == 1
    assert p.gamma_mode == 0 and p.chain_groups == 0
    assert ctypes.sizeof(capi.Params) == 8 * 4 + 8 + 8 + 6 * 8 + 4 * 4


def test_no_gpu_is_a_loud_error(bnr):
    """There is no CPU fallback: without a device bnr_create fails with BNR_ENODEV and a message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    X = np.zeros((4, 10))
    with pytest.raises(bnr.BnrError) as ei:
        bnr.Engine(X, np.zeros(4), 3)
    assert ei.value.code == -5 and "no CPU fallback" in str(ei.value)


def test_bad_arguments_rejected_before_any_cuda_call(bnr):
    from bnr_b200 import capi
    L = bnr.lib()
    assert L.bnr_create(None, None, None, None) == -1
    assert b"null" in L.bnr_last_error()
    with pytest.raises(ValueError):
        bnr.Engine(np.zeros((4, 11)), np.zeros(4), 3)        # 11 is not V(V+1)/2
    with pytest.raises(ValueError):
        bnr.Fit(np.zeros((4, 10)), np.zeros(4), 12, ν=10, x_transform=False, filename=None)   # nu < R


def test_summary_matches_goldens(bnr, golden):
    """Host-side Summary == the reference's stored `out2` table when fed the reference's stored traces."""
    st = bnr.Table(gamma=golden["res2.gamma"], xi=golden["res2.xi"])
    res = bnr.Results(st, golden["res2.rhat_xi"], golden["res2.rhat_gamma"], 200, 200)
    out = bnr.Summary(res)
    for k in ("node1", "node2", "estimate", "lower_bound", "upper_bound"):
        np.testing.assert_array_equal(out.edge_coef[k], golden["out2." + k], err_msg=k)
    np.testing.assert_array_equal(out.prob_nodes["probability"], golden["out2.probability"])
    assert "Edge Coefficient Estimates (95% credible intervals)" in repr(out)
    out90 = bnr.Summary(res, interval=90, digits=2)
    want = O.summary(golden["res2.gamma"][200:400, :, 0], golden["res2.xi"][200:400, :, 0], interval=90, digits=2)
    np.testing.assert_array_equal(out90.edge_coef["lower_bound"], want["lower_bound"])


def test_setup_X_and_index_maps(bnr):
    rng = np.random.default_rng(0)
    mats = []
    for _ in range(5):
        A = rng.random((6, 6))
        mats.append(A + A.T)
    Xn = bnr.setup_X(mats)
    assert Xn.shape == (5, 21)
    np.testing.assert_array_equal(Xn, O.setup_X(mats))
    np.testing.assert_array_equal(bnr.create_lower_tri(Xn[0], 6), np.tril(mats[0]))
    np.testing.assert_array_equal(bnr.lower_triangle(mats[1]), O.lower_triangle(mats[1]))
    with pytest.raises(ValueError):
        bnr.lower_triangle(np.zeros((2, 3)))


def test_table_aliases(bnr):
    t = bnr.Table(gamma=np.zeros((3, 2, 1)), xi=np.ones((3, 4, 1)), tau2=np.zeros((3, 1, 1)), pi=np.zeros((3, 2, 3)))
    assert t["γ"] is t["gamma"] is t.gamma is t.γ
    assert t["ξ"] is t.xi and t["τ²"] is t.tau2 and t["πᵥ"] is t.pi
    assert len(t) == 3


def test_input_formats_roundtrip(bnr, tmp_path, golden):
    """The documented input workflows (docs/src/man/inputdata.md): vectorised CSV with the response in the last
    column, and one CSV per adjacency matrix + responses.csv; both must give the same n x q design matrix."""
    X, y = golden["example.X"][:7], golden["example.y"][:7]
    V = 30
    vec = tmp_path / "matrix_networks.csv"
    hdr = ",".join("x%d" % (i + 1) for i in range(X.shape[1] + 1))
    np.savetxt(vec, np.column_stack([X, y]), delimiter=",", header=hdr, comments="")
    Xr, yr = bnr.read_matrix_networks(str(vec))
    np.testing.assert_allclose(Xr, X, rtol=1e-15)
    np.testing.assert_allclose(yr, y, rtol=1e-15)
    for i in range(len(y)):
        A = bnr.create_lower_tri(X[i], V)
        A = A + np.tril(A, -1).T
        np.savetxt(tmp_path / ("data%d.csv" % (i + 1)), A, delimiter=",",
                   header=",".join("Column%d" % (c + 1) for c in range(V)), comments="")
    np.savetxt(tmp_path / "responses.csv", y, header="Column1", comments="")
    mats = bnr.read_adjacency_csvs(str(tmp_path), len(y))
    assert len(mats) == len(y) and mats[0].shape == (V, V)
    np.testing.assert_allclose(bnr.setup_X(mats, True), X, rtol=1e-15)
    np.testing.assert_allclose(bnr.read_responses(str(tmp_path / "responses.csv")), y, rtol=1e-15)
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.csv"
        np.savetxt(bad, np.zeros((3, 8)), delimiter=",", header="a,b,c,d,e,f,g,h", comments="")
        bnr.read_matrix_networks(str(bad))


def _build_c_client(tmp_path, name="abi_smoke"):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "bayesiannetworkregression.jl_b200")
    exe = str(tmp_path / name)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "c", name + ".c"), "-o", exe, "-L", pkg, "-lbnr",
                    "-Wl,-rpath," + pkg, "-lm"], check=True)
    return exe


def test_header_is_plain_c_and_links(bnr, tmp_path):
    """include/bnr.h compiles as strict C99 and a plain-C client links against libbnr.so: the boundary a Julia ccall /
    cgo / JNI binding sees.  Without a GPU the client must stop with BNR_ENODEV (exit code 3), never fall back."""
    import subprocess
    import torch
    exe = _build_c_client(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


def test_c_fit_client_links_and_refuses_without_gpu(bnr, tmp_path):
    """tests/c/fit_client.c drives a whole doubling-scheme Fit through bnr_fit from plain C (on the GPU: see
    tests/test_gpu_engine.py); here: it compiles as strict C99 against the header and stops with BNR_ENODEV."""
    import subprocess
    import torch
    exe = _build_c_client(tmp_path, "fit_client")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3, r.stdout + r.stderr


def test_product_never_touches_the_oracle():
    """The oracle is test / bench infrastructure: nothing under the package may import, call or link it, and the
    package has no CPU fallback path (every numerical entry point goes through libbnr.so)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "bayesiannetworkregression.jl_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), encoding="utf8").read()
                for line in text.splitlines():
                    code = line.split("#")[0].split("//")[0]
                    if re.search(r"\b(import|from)\s+oracle\b|oracle\.|oracle/", code):
                        offenders.append((f, line.strip()))
    assert not offenders, offenders
    eng_src = open(os.path.join(pkg, "engine.py"), encoding="utf8").read()
    assert "numpy.linalg" not in eng_src and "np.linalg" not in eng_src and "scipy" not in eng_src
