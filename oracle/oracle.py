"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

Python face of the CPU oracle for the frontier log-likelihood path of elaine84/fetching:

* ctypes bindings over ``oracle/_build/liboracle.so`` (``loglik_oracle.c``, the plain-C
  restatement; every function there cites the reference file:line it follows);
* ``np_mixture_evaluate`` / ``np_blasso_evaluate`` / ``np_blasso_log_prior``: the numpy
  statements of pysamplers/pymixture.py:86-92 and pysamplers/pyblasso.py:123-141, i.e. what
  a reference worker really executes per item (python.cc:182-192);
* ``run_pf_ref``: runs ``oracle/_ref/pf_ref`` -- the reference's own C++ compiled in place
  against oracle/shim (see oracle/Makefile) -- when that binary is present.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  ``fetching_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "pf_ref")
REF_GPU_BIN = os.path.join(HERE, "_ref", "pf_ref_gpu")

_dp = C.POINTER(C.c_double)
_ldp = C.POINTER(C.c_longdouble)
_lib = None


def build(force: bool = False) -> None:
    """Compile the C restatement (and the in-place reference build when /root/reference exists)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(HERE, "loglik_oracle.c")):
        subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(REF_BIN)):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
    # the reference tree + the product's C ABI in one binary (needs the built libpfetch.so; runs on a GPU box only)
    pflib = os.path.join(os.path.dirname(HERE), "fetching_b200", "lib", "libpfetch.so")
    srcs = [os.path.join(HERE, f) for f in ("ref_gpu_driver.cc", "ref_driver.cc")] + [pflib]
    if os.path.isdir("/root/reference") and os.path.exists(pflib) and (
            force or not os.path.exists(REF_GPU_BIN) or
            os.path.getmtime(REF_GPU_BIN) < max(os.path.getmtime(x) for x in srcs)):
        subprocess.run(["make", "-s", "-C", HERE, "ref-gpu"], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        sz = C.c_size_t
        L.orc_mt_stream.argtypes = [C.c_uint32, sz, C.POINTER(C.c_uint32)]
        L.orc_normal_dataset.argtypes = [sz, _dp]
        L.orc_valid_stride.argtypes = [sz, sz]
        L.orc_valid_stride.restype = sz
        L.orc_stride_index.argtypes = [sz, sz, sz, sz]
        L.orc_stride_index.restype = sz
        L.orc_normal_prepare_draws.argtypes = [sz, sz, C.POINTER(sz), C.POINTER(sz)]
        L.orc_normal_prepare_draws.restype = sz
        L.orc_normal_propose.argtypes = [C.c_int, sz, _dp]
        L.orc_normal_propose.restype = sz
        L.orc_normal_eval.argtypes = [_dp, sz, C.c_int, _dp, sz, sz, sz, sz, _dp, _dp]
        L.orc_normal_eval_ld.argtypes = [_dp, sz, C.c_int, _dp, _ldp, _ldp]
        L.orc_boltzmann_scalar_eval.argtypes = [C.c_double, C.c_double, sz, _dp, _dp]
        L.orc_boltzmann_dense_eval.argtypes = [_dp, sz, _dp, C.c_double, sz, sz, _dp, _dp]
        L.orc_boltzmann_dense_eval_ld.argtypes = [_dp, sz, _dp, C.c_double, _ldp, _ldp]
        L.orc_mixture_item.argtypes = [_dp, sz, C.c_int, C.c_int, _dp, sz, sz]
        L.orc_mixture_item.restype = C.c_double
        L.orc_mixture_eval.argtypes = [_dp, sz, C.c_int, C.c_int, _dp, sz, sz, sz, _dp, _dp]
        L.orc_mixture_eval_ld.argtypes = [_dp, sz, C.c_int, C.c_int, _dp, sz, sz, _ldp, _ldp]
        L.orc_blasso_item.argtypes = [_dp, _dp, sz, C.c_int, _dp, sz, sz]
        L.orc_blasso_item.restype = C.c_double
        L.orc_blasso_eval.argtypes = [_dp, _dp, sz, C.c_int, _dp, sz, sz, sz, _dp, _dp]
        L.orc_blasso_eval_ld.argtypes = [_dp, _dp, sz, C.c_int, _dp, sz, sz, _ldp, _ldp]
        L.orc_blasso_log_prior.argtypes = [_dp, C.c_int, C.c_double]
        L.orc_blasso_log_prior.restype = C.c_double
        L.orc_mixture_frontier_mt.argtypes = [_dp, sz, C.c_int, C.c_int, _dp, sz, sz, sz, _dp, _dp]
        L.orc_blasso_frontier_mt.argtypes = [_dp, _dp, sz, C.c_int, _dp, sz, sz, sz, _dp, _dp]
        L.orc_normal_frontier_mt.argtypes = [_dp, sz, C.c_int, _dp, sz, _dp, _dp]
        L.orc_boltzmann_dense_logp.argtypes = [_dp, sz, _dp, C.c_double]
        L.orc_boltzmann_dense_logp.restype = C.c_double
        L.orc_boltzmann_dense_frontier_mt.argtypes = [_dp, sz, _dp, sz, C.c_double, _dp, _dp]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


# ----------------------------------------------------------------------------- RNG / data
def mt_stream(seed: int, count: int) -> np.ndarray:
    out = np.zeros(count, dtype=np.uint32)
    lib().orc_mt_stream(seed, count, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


def normal_dataset(n_points: int) -> np.ndarray:
    """samplers/normal.cc:41-46."""
    out = np.zeros((n_points, 2))
    lib().orc_normal_dataset(n_points, _p(out))
    return out


def valid_stride(n: int, stride: int) -> int:
    return int(lib().orc_valid_stride(n, stride))


def normal_prepare_draws(n: int, random_position: int):
    """(stride, start, next_random_position) drawn by NormalSampler::prepare."""
    s, o = C.c_size_t(), C.c_size_t()
    pos = lib().orc_normal_prepare_draws(n, random_position, C.byref(s), C.byref(o))
    return int(s.value), int(o.value), int(pos)


def normal_propose(theta: np.ndarray, random_position: int, D: int = 2):
    t = _c(theta).copy()
    pos = lib().orc_normal_propose(D, random_position, _p(t))
    return t, int(pos)


# ----------------------------------------------------------------------------- evaluators
def normal_eval(data, theta, stride=1, start=0, first=0, count=None, sum0=0.0, sq0=0.0):
    data, theta = _c(data), _c(theta)
    n, D = data.shape
    if count is None:
        count = n - first
    s, q = C.c_double(sum0), C.c_double(sq0)
    lib().orc_normal_eval(_p(data), n, D, _p(theta), stride, start, first, count, C.byref(s), C.byref(q))
    return s.value, q.value


def normal_eval_ld(data, theta):
    data, theta = _c(data), _c(theta)
    n, D = data.shape
    s, q = C.c_longdouble(0), C.c_longdouble(0)
    lib().orc_normal_eval_ld(_p(data), n, D, _p(theta), C.byref(s), C.byref(q))
    return np.longdouble(s.value), np.longdouble(q.value)


def boltzmann_scalar_eval(theta: float, count: int, temperature: float = 1.0):
    s, q = C.c_double(0), C.c_double(0)
    lib().orc_boltzmann_scalar_eval(theta, temperature, count, C.byref(s), C.byref(q))
    return s.value, q.value


def boltzmann_dense_eval(W, x, temperature=1.0, first=0, count=None):
    W, x = _c(W), _c(x)
    D = W.shape[0]
    if count is None:
        count = D - first
    s, q = C.c_double(0), C.c_double(0)
    lib().orc_boltzmann_dense_eval(_p(W), D, _p(x), temperature, first, count, C.byref(s), C.byref(q))
    return s.value, q.value


def boltzmann_dense_logp(W, x, temperature=1.0):
    """The single item value: -x'Wx/T."""
    W, x = _c(W), _c(x)
    return float(lib().orc_boltzmann_dense_logp(_p(W), W.shape[0], _p(x), temperature))


def boltzmann_dense_logp_independent(W, x, temperature=1.0):
    """A SECOND, unrelated statement of the dense Boltzmann energy -x'Wx/T, for a model that has no reference
    implementation to pin against (boltzmannsampler.hh:71-74 is a scalar toy): instead of the row-by-row double loop of
    loglik_oracle.c it sums the symmetric half
        sum_i W_ii x_i^2 + sum_{i<j} (W_ij + W_ji) x_i x_j
    in 80-bit extended precision with numpy's pairwise summation.  Different formula path, different summation order,
    different precision: agreement of the C restatement, the CUDA-core kernel and the DMMA kernel with THIS is the
    strongest check the workload admits (it remains "parity unpinned" with respect to the reference)."""
    W = np.asarray(W, dtype=np.longdouble)
    x = np.asarray(x, dtype=np.longdouble)
    iu = np.triu_indices(W.shape[0], k=1)
    diag = np.sum(np.diagonal(W) * x * x)
    off = np.sum((W[iu] + W.T[iu]) * (x[iu[0]] * x[iu[1]]))
    return -(diag + off) / np.longdouble(temperature)


def boltzmann_dense_eval_ld(W, x, temperature=1.0):
    W, x = _c(W), _c(x)
    s, q = C.c_longdouble(0), C.c_longdouble(0)
    lib().orc_boltzmann_dense_eval_ld(_p(W), W.shape[0], _p(x), temperature, C.byref(s), C.byref(q))
    return np.longdouble(s.value), np.longdouble(q.value)


def mixture_n_items(n_rows: int, item_rows: int) -> int:
    """pymixture.py:64-67 (Python 2 integer division)."""
    return (n_rows + 1) // item_rows


def mixture_item(data, theta, item_rows, position):
    data, theta = _c(data), _c(theta)
    n, d = data.shape
    K = theta.size // d
    return float(lib().orc_mixture_item(_p(data), n, d, K, _p(theta), item_rows, position))


def mixture_eval(data, theta, item_rows, first=0, count=None):
    data, theta = _c(data), _c(theta)
    n, d = data.shape
    K = theta.size // d
    if count is None:
        count = mixture_n_items(n, item_rows) - first
    s, q = C.c_double(0), C.c_double(0)
    lib().orc_mixture_eval(_p(data), n, d, K, _p(theta), item_rows, first, count, C.byref(s), C.byref(q))
    return s.value, q.value


def mixture_eval_ld(data, theta, item_rows):
    data, theta = _c(data), _c(theta)
    n, d = data.shape
    K = theta.size // d
    s, q = C.c_longdouble(0), C.c_longdouble(0)
    lib().orc_mixture_eval_ld(_p(data), n, d, K, _p(theta), item_rows, mixture_n_items(n, item_rows),
                              C.byref(s), C.byref(q))
    return np.longdouble(s.value), np.longdouble(q.value)


def blasso_n_items(n_rows: int, item_rows: int) -> int:
    """pyblasso.py:68-69."""
    return n_rows // item_rows


def blasso_item(X, y, theta, item_rows, position):
    X, y, theta = _c(X), _c(y), _c(theta)
    n, p = X.shape
    return float(lib().orc_blasso_item(_p(X), _p(y), n, p, _p(theta), item_rows, position))


def blasso_eval(X, y, theta, item_rows, first=0, count=None):
    X, y, theta = _c(X), _c(y), _c(theta)
    n, p = X.shape
    if count is None:
        count = blasso_n_items(n, item_rows) - first
    s, q = C.c_double(0), C.c_double(0)
    lib().orc_blasso_eval(_p(X), _p(y), n, p, _p(theta), item_rows, first, count, C.byref(s), C.byref(q))
    return s.value, q.value


def blasso_eval_ld(X, y, theta, item_rows):
    X, y, theta = _c(X), _c(y), _c(theta)
    n, p = X.shape
    s, q = C.c_longdouble(0), C.c_longdouble(0)
    lib().orc_blasso_eval_ld(_p(X), _p(y), n, p, _p(theta), item_rows, blasso_n_items(n, item_rows),
                             C.byref(s), C.byref(q))
    return np.longdouble(s.value), np.longdouble(q.value)


def blasso_log_prior(theta, tau=5.0):
    theta = _c(theta)
    return float(lib().orc_blasso_log_prior(_p(theta), theta.size - 1, tau))


# ----------------------------------------------------------------------------- frontier (all host threads)
def max_threads() -> int:
    return int(lib().orc_max_threads())


def set_threads(n: int) -> int:
    """Use exactly `n` OpenMP threads from now on, whatever OMP_NUM_THREADS says; returns the count in force."""
    return int(lib().orc_set_threads(int(n)))


def mixture_frontier(data, thetas, item_rows):
    data, thetas = _c(data), _c(thetas)
    n, d = data.shape
    B = thetas.shape[0]
    K = thetas[0].size // d
    s, q = np.zeros(B), np.zeros(B)
    lib().orc_mixture_frontier_mt(_p(data), n, d, K, _p(thetas), B, item_rows, mixture_n_items(n, item_rows),
                                  _p(s), _p(q))
    return s, q


def blasso_frontier(X, y, thetas, item_rows):
    X, y, thetas = _c(X), _c(y), _c(thetas)
    n, p = X.shape
    B = thetas.shape[0]
    s, q = np.zeros(B), np.zeros(B)
    lib().orc_blasso_frontier_mt(_p(X), _p(y), n, p, _p(thetas), B, item_rows, blasso_n_items(n, item_rows),
                                 _p(s), _p(q))
    return s, q


def normal_frontier(data, thetas):
    data, thetas = _c(data), _c(thetas)
    n, D = data.shape
    B = thetas.shape[0]
    s, q = np.zeros(B), np.zeros(B)
    lib().orc_normal_frontier_mt(_p(data), n, D, _p(thetas), B, _p(s), _p(q))
    return s, q


def boltzmann_dense_frontier(W, X, temperature=1.0):
    W, X = _c(W), _c(X)
    B = X.shape[0]
    s, q = np.zeros(B), np.zeros(B)
    lib().orc_boltzmann_dense_frontier_mt(_p(W), W.shape[0], _p(X), B, temperature, _p(s), _p(q))
    return s, q


# ----------------------------------------------------------------------------- numpy statements
def np_mixture_evaluate(data, theta, batch, position):
    """pymixture.py:86-92, statement for statement (theta is [K][d])."""
    lp = np.zeros(len(data[(position * batch):((position + 1) * batch)]))
    for i in range(theta.shape[0]):
        td = theta[i] - data[(position * batch):((position + 1) * batch)]
        lp += np.exp(-(td ** 2) / 2.).prod(axis=1)
    return np.log(lp).sum()


def np_blasso_evaluate(X_all, y_all, theta, position, item_rows=18000):
    """pyblasso.py:135-141, statement for statement."""
    X = X_all[(position * item_rows):((position + 1) * item_rows)]
    y = y_all[(position * item_rows):((position + 1) * item_rows)]
    sigma = theta[0]
    beta = theta[1:]
    lp = (np.log(1 / sigma / np.sqrt(2 * np.pi)) - (np.dot(beta, X.T) - y) ** 2 / 2 / sigma ** 2).sum()
    return lp


def np_blasso_log_prior(theta, ndim, tau=5.0):
    """pyblasso.py:123-133."""
    sigma = theta[0]
    beta = theta[1:]
    log_prior_sigma = -2.0 * np.log(sigma)
    log_laplace = (ndim * np.log(tau / 2 / sigma) - (tau * np.abs(beta).sum()) / sigma)
    return log_prior_sigma + log_laplace


# ----------------------------------------------------------------------------- the reference binary
def have_pf_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_pf_ref(*args: str) -> str:
    return subprocess.run([REF_BIN, *map(str, args)], check=True, capture_output=True, text=True).stdout


def pf_ref_normal_eval(data: np.ndarray, thetas: np.ndarray) -> np.ndarray:
    """Reference NormalSampler<2>::evaluate_many over the whole dataset, per node -> [B][2]."""
    data, thetas = _c(data), _c(thetas)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<QQQ", data.shape[0], thetas.shape[0], thetas.shape[1]))
            f.write(data.tobytes())
            f.write(thetas.tobytes())
        subprocess.run([REF_BIN, "normal-eval", fin, fout], check=True)
        return np.fromfile(fout, dtype=np.float64).reshape(-1, 2)


def pf_ref_boltz_eval(thetas: np.ndarray, count: int) -> np.ndarray:
    thetas = _c(thetas)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<QQQ", count, thetas.size, 1))
            f.write(thetas.tobytes())
        subprocess.run([REF_BIN, "boltz-eval", fin, fout], check=True)
        return np.fromfile(fout, dtype=np.float64).reshape(-1, 2)


def parse_chain(text: str):
    """Rows of `pf_ref chain` / tests/golden/chain_*.tsv -> list of (depth, accept, metric_sum, theta_bytes)."""
    rows = []
    for line in text.splitlines():
        if not line or line.startswith("#"):
            continue
        f = line.split("\t")
        rows.append((int(f[0]), int(f[1]), float(f[2]), bytes.fromhex(f[3])))
    return rows
