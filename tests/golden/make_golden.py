"""Extract the reference's own golden vectors into a small fixture that can travel to the GPU box.

Run HERE (where /root/reference is mounted):  python tests/golden/make_golden.py
Sources (never copied as files, only decoded to arrays):
  /root/reference/test/data/gen_test_results.jld2   keys st1 st2 st3 res out res2 out2
      (written by test/gen_tst_results.jl:237, consumed by test/init-tests.jl:64-124,
       test/toy-generate-samples-test.jl:23-54, test/test1-generate-samples-test.jl:15-45)
  /root/reference/test/data/test1.csv               70 x 190 predictors + y (the data res2 was fit on)
  /root/reference/examples/true_xi.csv, true_b.csv  ground truth of the shipped example
Output: tests/golden/golden.npz (float64/int64 arrays, Julia dims kept, names "<key>.<field>").
"""
import os
import struct
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
from jld2_reader import JLD2File, read_table, STATE_FIELDS  # noqa: E402

REF = "/root/reference"
LIVE = STATE_FIELDS[:11]  # Sigma_inv / invC / mu_t are never written by the reference (garbage)


def dataframe_columns(jf, addr):
    raw = jf.dataset(addr)[2]
    cols_ref = struct.unpack_from("<Q", raw, 0)[0]
    craw = jf.dataset(cols_ref)[2]
    return [jf.dataset(a)[2] for a in struct.unpack("<%dQ" % (len(craw) // 8), craw)]


def main():
    jf = JLD2File(os.path.join(REF, "test/data/gen_test_results.jld2"))
    links = jf.links(jf.root)
    out = {}
    for key in ("st1", "st2", "st3"):
        cols = jf.refs(links[key])
        for name, a in zip(LIVE, cols):
            out["%s.%s" % (key, name)] = jf.f64(a)
    for key in ("res", "res2"):
        r = jf.refs(links[key])
        tab = read_table(jf, r[0])
        for name in LIVE:
            out["%s.%s" % (key, name)] = tab[name]
        out[key + ".rhat_xi"] = jf.f64(jf.refs(r[1])[0])
        out[key + ".rhat_gamma"] = jf.f64(jf.refs(r[2])[0])
        out[key + ".burn_in"] = np.int64(r[3])
        out[key + ".sampled"] = np.int64(r[4])
    for key in ("out", "out2"):
        r = jf.refs(links[key])
        ec = dataframe_columns(jf, r[0])
        out[key + ".node1"] = np.frombuffer(ec[0], "<i8").copy()
        out[key + ".node2"] = np.frombuffer(ec[1], "<i8").copy()
        out[key + ".estimate"] = np.frombuffer(ec[2], "<f8").copy()
        out[key + ".lower_bound"] = np.frombuffer(ec[3], "<f8").copy()
        out[key + ".upper_bound"] = np.frombuffer(ec[4], "<f8").copy()
        pn = dataframe_columns(jf, r[1])
        out[key + ".probability"] = np.frombuffer(pn[0], "<f8").copy()
        out[key + ".ci_level"] = np.int64(r[2])
    t1 = np.loadtxt(os.path.join(REF, "test/data/test1.csv"), delimiter=",", skiprows=1)
    out["test1.X"] = t1[:, :190].copy()
    out["test1.y"] = t1[:, 190].copy()
    # BASELINE config 1 input: the shipped vectorised example (100 x 465 lower triangles + response in the last
    # column, docs/src/man/inputdata.md:83-91) and the posterior tables of the reference's own stored 50 000-iteration
    # fit of the same data (older diagonal-free model, R = 7): soft level-2 references
    ex = np.loadtxt(os.path.join(REF, "examples/matrix_networks.csv"), delimiter=",", skiprows=1)
    out["example.X"] = ex[:, :465].copy()
    out["example.y"] = ex[:, 465].copy()
    tdir = os.path.join(REF, "test/data")
    pre = "R=7_mu=1.6_n_microbes=22_nu=10_out="
    suf = "_pi=0.8_samplesize=100_simnum=1.csv"
    ed = np.genfromtxt(os.path.join(tdir, pre + "edges" + suf), delimiter=",", names=True)
    nd = np.genfromtxt(os.path.join(tdir, pre + "nodes" + suf), delimiter=",", names=True)
    out["example.ref_edge_mean"] = np.asarray(ed["mean"], dtype=np.float64)
    out["example.ref_edge_lo"] = np.asarray(ed["0025"], dtype=np.float64)
    out["example.ref_edge_hi"] = np.asarray(ed["0975"], dtype=np.float64)
    out["example.ref_xi_posterior"] = np.asarray(nd["Xi_posterior"], dtype=np.float64)
    out["example.true_xi"] = _csv_col(os.path.join(REF, "examples/true_xi.csv"))
    out["example.true_b"] = _csv_col(os.path.join(REF, "examples/true_b.csv"))
    dst = os.path.join(os.path.dirname(__file__), "golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes;", len(out), "arrays")


def _csv_col(path):
    vals = []
    with open(path) as fh:
        for i, line in enumerate(fh):
            tok = line.strip().split(",")[-1].strip('"')
            try:
                vals.append(float({"true": 1, "false": 0}.get(tok.lower(), tok)))
            except ValueError:
                if i:
                    raise
    return np.asarray(vals)


if __name__ == "__main__":
    main()
