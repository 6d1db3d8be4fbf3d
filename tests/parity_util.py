"""Tolerance rule and bookkeeping shared by the GPU parity tests.

north_star: conditional parameters, means, Cholesky factors and draws agree with the reference to 1e-10 relative in
FP64.  A quantity x that comes out of a linear solve cannot be held to a fixed 1e-10 when the system is ill conditioned:
reference (LU), oracle (LAPACK) and engine (blocked Cholesky) are all only backward stable, |dx| / |x| <= c * cond * eps.
Measured against a long-double arbiter (oracle/extended.py, tests/test_gpu_parity_edges.py::
test_solve_error_follows_cond_times_eps) the constant is c < 0.6 for the engine and c < 2.2 for the float64 oracle, for
cond from 3e1 to 1e9; the rule below uses c = 8."""
import json
import os

import numpy as np

EPS = float(np.finfo(np.float64).eps)
C_SOLVE = 8.0


def solve_tol(cond):
    return max(1e-10, C_SOLVE * float(cond) * EPS)


def assert_close_normwise(got, want, tol, msg=""):
    """|got - want| <= tol * max|want| element by element (the norm-wise bound backward stability gives) and, for the
    entries that are not tiny against the largest one, the usual relative test as well."""
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    scale = float(np.max(np.abs(want))) if want.size else 1.0
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol * scale, err_msg=msg)


def rel_err(got, want):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    scale = float(np.max(np.abs(want))) if want.size else 1.0
    return float(np.max(np.abs(got - want)) / scale) if want.size and scale > 0 else 0.0


def note(**kw):
    """Append a measurement to gpurun_out/parity_constants.jsonl (scratch evidence; summarised under profiles/)."""
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_constants.jsonl"), "a") as fh:
            fh.write(json.dumps(kw) + "\n")
