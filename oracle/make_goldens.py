#!/usr/bin/env python
"""oracle/make_goldens.py -- TEST INFRASTRUCTURE ONLY; run in the authoring container.

Produces the committed fixtures under tests/golden/ from the REFERENCE ITSELF
(/root/reference, read-only, not present on the GPU box):

  chain_*.tsv            `oracle/_ref/pf_ref chain ...`: the reference JobTree + run_master +
                         MasterTester + NormalSampler<2>/BoltzmannSampler, compiled in place
                         against oracle/shim (oracle/Makefile).  One row per chain level.
  ref_normal_eval.npz    reference NormalSampler<2>::load/prepare/evaluate_many
                         (normalsampler.hh:66-165) on the normal.cc:41-46 dataset for a set
                         of thetas, including general (non-identity) covariances.
  ref_boltz_eval.npz     reference BoltzmannSampler::evaluate_many (boltzmannsampler.hh:71-82).
  py_models.npz          outputs of the reference's own pysamplers/pymixture.py and
                         pysamplers/pyblasso.py classes.  Those files are Python 2; they are
                         read from /root/reference at generation time, given the minimal
                         textual Python-3 patches listed in PY3_PATCHES (each asserted to
                         apply, so drift is detected), exec'd, and their methods
                         (Mixture.evaluate / log_posterior / first_proposal / data_size,
                         BLasso.evaluate / log_prior / log_contributions / data_size) called.
                         pyblasso reads external OPV .npy files that are not in the reference
                         repo (pyblasso.py:37-45); synthetic stand-ins are written to a temp
                         directory laid out the way it expects.

Usage:  python oracle/make_goldens.py        (writes tests/golden/*, prints a summary)
"""
import hashlib
import io
import contextlib
import os
import subprocess
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

PY3_PATCHES = {
    "pymixture.py": [
        ("import numpy as np\n", "np = _np_compat\n"),
        ("            print 'Did not use speculative rand.random()'\n",
         "            print('Did not use speculative rand.random()')\n"),
        ("        self.num_batches_ = (train_size + 1) / batch_size\n",
         "        self.num_batches_ = (train_size + 1) // batch_size\n"),
        ("    batch_size = train_size / 100\n", "    batch_size = train_size // 100\n"),
    ],
    "pyblasso.py": [
        ("import cPickle\n", "import pickle as cPickle\n"),
        ("import numpy as np\n", "np = _np_compat\n"),
        ("            print 'Did not use speculative rand.random()'\n",
         "            print('Did not use speculative rand.random()')\n"),
        ("            print 'Keeping %d / %d columns' % (len(cols), X.shape[1])\n",
         "            print('Keeping %d / %d columns' % (len(cols), X.shape[1]))\n"),
        ("        self.data_size_ = self.data_size_ / 18000\n",
         "        self.data_size_ = self.data_size_ // 18000\n"),
    ],
}


class _NpCompat(types.ModuleType):
    """numpy plus the `np.cast[int](x)` spelling removed in numpy 2.0 (pymixture.py:58)."""

    def __init__(self):
        super().__init__("numpy_compat")
        self.cast = {int: lambda a: np.asarray(a).astype(int)}

    def __getattr__(self, name):
        return getattr(np, name)


def load_reference_module(fname):
    """Exec the class/function definitions of a reference Python-2 model under Python 3."""
    src = open(os.path.join(REF, "pysamplers", fname)).read()
    for old, new in PY3_PATCHES[fname]:
        assert old in src, "reference source drifted: %r not in %s" % (old, fname)
        src = src.replace(old, new)
    # Everything from the first serial-driver helper on (`def sequential` .. end of class, plus
    # `__main__`) is plotting / printing code in Python-2 syntax and is off the evaluation path:
    # cut each class at `    def sequential` / `    def check` and keep the module-level `sampler`.
    keep = []
    skipping = False
    for line in src.splitlines(True):
        if line.startswith("    def sequential") or line.startswith("    def check"):
            skipping = True
        elif skipping and (line.startswith("def ") or line.startswith("class ")):
            skipping = False
        if line.startswith("if __name__"):
            break
        if not skipping:
            keep.append(line)
    ns = {"_np_compat": _NpCompat(), "__name__": "reference_" + fname[:-3]}
    exec(compile("".join(keep), os.path.join(REF, "pysamplers", fname), "exec"), ns)
    return ns


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gen_chains():
    runs = {
        "chain_normal_d10000.tsv": ["chain", "normal", "-n", "9", "-d", "10000", "-b", "1000", "-l", "300"],
        "chain_normal_d100000.tsv": ["chain", "normal", "-n", "17", "-d", "100000", "-b", "10000", "-l", "60"],
        "chain_boltzmann_d10000.tsv": ["chain", "boltzmann", "-n", "9", "-d", "10000", "-b", "1000", "-l", "300"],
    }
    for name, args in runs.items():
        out = orc.run_pf_ref(*args)
        # the chain must not depend on how much was speculated: re-run with
        # different simulated core counts / batch sizes / policies and insist on identical text
        # (depth, accept, theta bytes) bit-identical; metric_sum to 1e-12 because the partial sums
        # are accumulated per batch (worker.hh:148-155), so its last bits follow the batch size.
        base = orc.parse_chain(out)
        for extra in (["-n", "2"], ["-n", "65", "-b", "2500"], ["-n", "33", "-s", "0"]):
            other = orc.parse_chain(orc.run_pf_ref(*(args + extra)))
            assert [(r[0], r[1], r[3]) for r in other] == [(r[0], r[1], r[3]) for r in base], (name, extra)
            assert all(abs(a[2] - b[2]) <= 1e-12 * abs(b[2]) for a, b in zip(other, base)), (name, extra)
        with open(os.path.join(GOLD, name), "w") as f:
            f.write("# generated by oracle/make_goldens.py: pf_ref %s\n" % " ".join(args))
            f.write("# columns: depth accept metric_sum(%.17g) theta(hex of save_theta bytes) unparse(theta)\n")
            f.write(out)
        print(name, len(out.splitlines()), "rows")


FRONTIER_TRACES = {
    # file: pf_ref arguments -- the REFERENCE JobTree/run_master/samplers driven by the product's
    # FrontierCommunicator header (oracle/ref_driver.cc, `frontier`)
    "frontier_trace_normal_W8.tsv": ["frontier", "normal", "-W", "8", "-d", "10000", "-l", "80"],
    "frontier_trace_normal_W16_deferred_naive.tsv": ["frontier", "normal", "-W", "16", "-d", "10000", "-l", "60", "-s", "0",
                                                     "--deferred"],
    "frontier_trace_normal_W1.tsv": ["frontier", "normal", "-W", "1", "-d", "10000", "-l", "30"],
    "frontier_trace_normal_W64_const.tsv": ["frontier", "normal", "-W", "64", "-d", "10000", "-l", "60", "-s", "1"],
    "frontier_trace_boltzmann_W8.tsv": ["frontier", "boltzmann", "-W", "8", "-d", "10000", "-l", "80"],
    "frontier_trace_boltzmann_W5_deferred.tsv": ["frontier", "boltzmann", "-W", "5", "-d", "10000", "-l", "60", "--deferred"],
    # batched mode: partial updates feed the predictive scheduler, workers get abandoned (worker.hh:141-175)
    "frontier_trace_normal_W8_b1000_fastslow.tsv": ["frontier", "normal", "-W", "8", "-d", "10000", "-l", "80", "-b", "1000", "-y", "0"],
    "frontier_trace_normal_W16_b2000_incremental.tsv": ["frontier", "normal", "-W", "16", "-d", "10000", "-l", "40", "-b", "2000", "-y", "1"],
    "frontier_trace_boltzmann_W8_b1000_fastslow.tsv": ["frontier", "boltzmann", "-W", "8", "-d", "10000", "-l", "60", "-b", "1000", "-y", "0"],
}


def gen_frontier_traces():
    for name, args in FRONTIER_TRACES.items():
        out = subprocess.run([orc.REF_BIN, *args], check=True, capture_output=True, text=True).stdout
        rows = [l for l in out.splitlines() if l.startswith("R")]
        # the chain itself must be the one MasterTester produces (chain_*.tsv), whatever the frontier
        gold = orc.parse_chain(open(os.path.join(GOLD, "chain_%s_d10000.tsv" % args[1])).read())
        for r, g in zip(rows, gold):
            f = r.split("\t")
            assert (int(f[1]), int(f[2]), bytes.fromhex(f[4])) == (g[0], g[1], g[3]), (name, r)
            assert abs(float(f[3]) - g[2]) <= 1e-12 * abs(g[2])
        with open(os.path.join(GOLD, name), "w") as f:
            f.write("# generated by oracle/make_goldens.py: pf_ref %s\n" % " ".join(args))
            f.write("# J id generation state path random_position proposal_random_position next_random_position "
                    "log_prior theta(hex)\n# R depth accept metric_sum theta(hex)\n")
            f.write(out)
        print(name, len(out.splitlines()), "lines")


def gen_ref_normal():
    n = 10000
    data = orc.normal_dataset(n)
    ref_head = [float(x) for x in orc.run_pf_ref("data", 4).split()]
    assert list(data[:4].ravel()) == ref_head, "oracle dataset recipe != reference binary"
    rng = np.random.RandomState(1234)
    thetas = []
    for i in range(12):  # identity covariance (the only branch the reference chain ever takes)
        thetas.append([rng.normal(200 * (i % 2), 3.0), rng.normal(100 * (i % 2), 3.0), 1.0, 0.0, 1.0])
    thetas[0] = [0.0, 0.0, 1.0, 0.0, 1.0]  # the initial theta, normaltheta.hh:55-62
    for i in range(8):   # general covariance: packed lower triangle (v00, v10, v11)
        a, b = rng.uniform(0.5, 3.0, 2)
        c = rng.uniform(-0.9, 0.9) * np.sqrt(a * b)
        thetas.append([rng.normal(200, 1.0), rng.normal(100, 1.0), a, c, b])
    thetas = np.array(thetas)
    res = orc.pf_ref_normal_eval(data, thetas)
    np.savez(os.path.join(GOLD, "ref_normal_eval.npz"), n=n, thetas=thetas, sum=res[:, 0], sum_sq=res[:, 1],
             data_sha256=sha(data), data_head=data[:4].ravel())
    print("ref_normal_eval.npz", thetas.shape, res[0])

    th = np.array([0.25, 0.0, 1.5, 0.19217362052255774, 3.0e-7, 123.456])
    res = orc.pf_ref_boltz_eval(th, 10000)
    np.savez(os.path.join(GOLD, "ref_boltz_eval.npz"), thetas=th, count=10000, sum=res[:, 0], sum_sq=res[:, 1])
    print("ref_boltz_eval.npz", res[0])


def gen_py_models():
    out = {}
    # ---- Gaussian mixture, reference shape K=8, d=8 (pymixture.py:38-55)
    mod = load_reference_module("pymixture.py")
    with contextlib.redirect_stdout(io.StringIO()):
        m = mod["sampler"](["train_size=4000", "batch_size=250", "seed=0"])
        short = mod["sampler"](["train_size=3999", "batch_size=250", "seed=0"])
    assert m.data.shape == (4000, 8) and m.data_size() == 16
    # Corner: train_size = 3999 -> data_size() is still (3999+1)//250 = 16 (pymixture.py:64), but the
    # last item has 249 rows and `lp = zeros(batch)` cannot broadcast (pymixture.py:88-91): the
    # reference raises, and python.cc:188-191 turns that into a 0.0 contribution.
    assert short.data_size() == 16
    try:
        short.evaluate(short.first_proposal(), 15)
        raise SystemExit("expected the reference to fail on a short item")
    except ValueError:
        pass
    rng = np.random.RandomState(7)
    thetas = [m.first_proposal(), m.positions.copy()]
    for s in (0.05, 0.5, 2.0):
        thetas.append(m.positions[rng.permutation(8)] + rng.normal(0, s, (8, 8)))
    thetas = np.array(thetas)
    ev = np.array([[m.evaluate(t, pos) for pos in range(m.data_size())] for t in thetas])
    out.update(mix_data=m.data, mix_item_rows=250, mix_n_items=m.data_size(), mix_thetas=thetas, mix_item_lp=ev,
               mix_log_posterior=np.array([m.log_posterior(t) for t in thetas]),
               mix_positions=m.positions)
    print("mixture", m.data.shape, "items", m.data_size(), ev[0][:2])

    # ---- Bayesian lasso.  External OPV data is absent: synthesise files where it looks for them.
    mod = load_reference_module("pyblasso.py")
    p_raw, train, test = 5, 36000, 500
    rng = np.random.RandomState(11)
    Xraw = rng.normal(0, 1, (train + test, p_raw)) * rng.uniform(0.5, 4, p_raw) + rng.uniform(-2, 2, p_raw)
    Xraw[:, 3] = 1.0  # a constant column: dropped by the std > 0.001 filter (pyblasso.py:55)
    beta_true = np.array([0.8, -1.2, 0.0, 0.0, 2.5])
    yraw = Xraw @ beta_true + rng.normal(0, 0.7, train + test) + 3.0
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "data", "opv"))
        os.makedirs(os.path.join(td, "a", "b"))
        np.save(os.path.join(td, "data", "opv", "opv-response.npy"), yraw)
        np.save(os.path.join(td, "data", "opv", "opv-features.npy"), Xraw)
        cwd = os.getcwd()
        os.chdir(os.path.join(td, "a", "b"))
        try:
            np.random.seed(5)  # `rand` is undefined outside the C++ bridge -> bare except -> global numpy stream
            with contextlib.redirect_stdout(io.StringIO()):
                bl = mod["BLasso"](train_size=train, test_size=test, tau=5.0, quiet=True, seed=0, jump=None)
        finally:
            os.chdir(cwd)
    assert bl.data.shape == (train, 4) and bl.data_size() == 2
    rng = np.random.RandomState(13)
    thetas = []
    for sg in (0.2, 0.7, 1.0, 3.0):
        thetas.append(np.concatenate([[sg], rng.normal(0, 1.0, bl.ndim)]))
    thetas.append(np.concatenate([[0.69], [0.8, -1.2, 0.0, 2.5] * bl.data.std(axis=0)]))
    thetas = np.array(thetas)
    ev = np.array([[bl.evaluate(t, pos) for pos in range(bl.data_size())] for t in thetas])
    out.update(bl_X=bl.data, bl_y=bl.response, bl_item_rows=18000, bl_n_items=bl.data_size(), bl_thetas=thetas,
               bl_item_lp=ev, bl_log_prior=np.array([bl.log_prior(t) for t in thetas]),
               bl_contrib_sum=np.array([bl.log_contributions(t).sum() for t in thetas]), bl_tau=bl.tau)
    print("blasso", bl.data.shape, "items", bl.data_size(), ev[0], out["bl_log_prior"][0])
    out["numpy_version"] = np.array(np.__version__)
    np.savez_compressed(os.path.join(GOLD, "py_models.npz"), **out)


if __name__ == "__main__":
    assert os.path.isdir(REF), "needs /root/reference (authoring container only)"
    os.makedirs(GOLD, exist_ok=True)
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)
    gen_chains()
    gen_frontier_traces()
    gen_ref_normal()
    gen_py_models()
    for f in sorted(os.listdir(GOLD)):
        print("%10d  %s" % (os.path.getsize(os.path.join(GOLD, f)), f))
