"""Python face of the frontier log-likelihood engine.

`FrontierEngine` mirrors what a reference worker holds (the data set and an Executor,
samplers/normal.cc:38-49, samplers/python.cc:96-139) and what it returns per node
(`metric_sum`, `metric_sum_sq`, protocol.hh:43-53) -- for a whole frontier of nodes per call.
All arithmetic runs in libpfetch.so on the GPU; numpy is only the container for host buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import (F32, F64, MODEL_BLASSO, MODEL_BOLTZMANN_DENSE, MODEL_BOLTZMANN_SCALAR, MODEL_MIXTURE,
                   MODEL_NORMAL, ModelDesc, PfetchError, check)

_dp = C.POINTER(C.c_double)


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class FrontierEngine:
    """One engine = one device-resident data set (or row shard) + the kernels of one model."""

    def __init__(self, model: int, data=None, *, response=None, dim: int = 0, components: int = 0,
                 item_rows: int = 0, n_items: int = 0, n_rows: Optional[int] = None, param: float = 0.0,
                 dtype: int = F64, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        desc = ModelDesc()
        desc.model, desc.dtype, desc.device = model, dtype, device
        keep = []
        if data is not None:
            data = _c(data)
            keep.append(data)
            if data.ndim == 1:
                data = data.reshape(-1, 1)
            desc.data = data.ctypes.data_as(_dp)
            desc.n_rows = data.shape[0] if n_rows is None else n_rows
            desc.dim = dim or data.shape[1]
        else:
            desc.n_rows = n_rows or 0
            desc.dim = dim
        if response is not None:
            response = _c(response)
            keep.append(response)
            desc.response = response.ctypes.data_as(_dp)
        desc.components = components
        desc.item_rows = item_rows
        desc.n_items = n_items
        desc.param = param
        check(self._lib.pf_engine_create(C.byref(desc), C.byref(self._h)), "pf_engine_create")
        self.model, self.dtype, self.device = model, dtype, device
        self.theta_len = int(self._lib.pf_engine_theta_len(self._h))
        self.data_size = int(self._lib.pf_engine_data_size(self._h))
        self.max_frontier = int(self._lib.pf_engine_max_frontier(self._h))

    # ---- constructors named after the reference's samplers
    @classmethod
    def normal(cls, data, **kw):
        """samplers/normal.cc + normalsampler.hh: Gaussian target over points data[n][D]."""
        return cls(MODEL_NORMAL, data, **kw)

    @classmethod
    def mixture(cls, data, components: int, batch_size: int = 0, **kw):
        """pysamplers/pymixture.py: K spherical unit-variance components, items of `batch_size` rows."""
        return cls(MODEL_MIXTURE, data, components=components, item_rows=batch_size, **kw)

    @classmethod
    def blasso(cls, X, y, item_rows: int = 18000, **kw):
        """pysamplers/pyblasso.py: theta = (sigma, beta), items of 18000 rows."""
        return cls(MODEL_BLASSO, X, response=y, item_rows=item_rows, **kw)

    @classmethod
    def boltzmann_dense(cls, W, temperature: float = 1.0, **kw):
        """BASELINE config 2: log p(x) = -x'Wx/T (new workload, one item)."""
        return cls(MODEL_BOLTZMANN_DENSE, W, param=temperature, **kw)

    @classmethod
    def boltzmann_scalar(cls, data_size: int, temperature: float = 1.0, **kw):
        """boltzmannsampler.hh: log p = -theta/T per item."""
        return cls(MODEL_BOLTZMANN_SCALAR, None, n_rows=data_size, dim=1, param=temperature, **kw)

    # ---- evaluation
    def _thetas(self, thetas) -> np.ndarray:
        t = _c(thetas)
        t = t.reshape(-1, self.theta_len) if t.ndim != 1 or self.theta_len != 1 else t.reshape(-1, 1)
        return t

    def eval_frontier(self, thetas) -> Tuple[np.ndarray, np.ndarray]:
        """(metric_sum[B], metric_sum_sq[B]) over all items; host buffers in, host buffers out."""
        t = self._thetas(thetas)
        B = t.shape[0]
        s, q = np.empty(B), np.empty(B)
        check(self._lib.pf_eval_frontier(self._h, B, t.ctypes.data_as(_dp), s.ctypes.data_as(_dp),
                                         q.ctypes.data_as(_dp)), "pf_eval_frontier")
        return s, q

    def theta_buffer(self) -> np.ndarray:
        """The engine's pinned staging buffer as a [max_frontier][theta_len] array (pf_engine_theta_buffer): a frontier
        written here and passed to the eval calls is DMA'd in place, without the engine's own staging copy."""
        cap = C.c_uint64(0)
        p = self._lib.pf_engine_theta_buffer(self._h, C.byref(cap))
        return np.ctypeslib.as_array(p, shape=(int(cap.value) // self.theta_len, self.theta_len))

    def eval_frontier_repeat(self, thetas, iters: int, pinned: bool = False):
        """`iters` complete host-buffer calls in a C loop; returns (sum, sum_sq, seconds).  pinned = True: the frontier
        is first placed in the engine's pinned staging buffer (outside the timed loop) and read from there."""
        t = self._thetas(thetas)
        B = t.shape[0]
        if pinned and B <= self.max_frontier:
            buf = self.theta_buffer()
            buf[:B] = t
            t = buf[:B]
        s, q = np.empty(B), np.empty(B)
        sec = C.c_double(0.0)
        check(self._lib.pf_eval_frontier_repeat(self._h, B, t.ctypes.data_as(_dp), s.ctypes.data_as(_dp),
                                                q.ctypes.data_as(_dp), iters, C.byref(sec)), "pf_eval_frontier_repeat")
        return s, q, float(sec.value)

    def eval_frontier_range(self, thetas, first_item: int, n_items: int, order=None):
        t = self._thetas(thetas)
        B = t.shape[0]
        s, q = np.empty(B), np.empty(B)
        op = None
        if order is not None:
            order = np.ascontiguousarray(order, dtype=np.uint64)
            op = order.ctypes.data_as(C.POINTER(C.c_uint64))
        check(self._lib.pf_eval_frontier_range(self._h, B, t.ctypes.data_as(_dp), first_item, n_items, op,
                                               s.ctypes.data_as(_dp), q.ctypes.data_as(_dp)),
              "pf_eval_frontier_range")
        return s, q

    def eval_frontier_device(self, d_thetas_ptr: int, d_out_ptr: int, B: int, stream_ptr: int = 0) -> None:
        """Device pointers (e.g. torch tensors' data_ptr()), enqueued on `stream_ptr`; no sync."""
        check(self._lib.pf_eval_frontier_device(self._h, B, C.c_void_p(d_thetas_ptr), C.c_void_p(d_out_ptr),
                                                C.c_void_p(stream_ptr)), "pf_eval_frontier_device")

    def set_option(self, option: int, value: int) -> None:
        check(self._lib.pf_engine_set_option(self._h, option, value), "pf_engine_set_option")

    def comm_init(self, unique_id: bytes, rank: int, nranks: int) -> None:
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(self._lib.pf_engine_comm_init(self._h, buf, rank, nranks), "pf_engine_comm_init")

    def peer_export(self) -> bytes:
        """64-byte CUDA IPC handle of this engine's exchange buffer (pf_engine_peer_export)."""
        buf = C.create_string_buffer(64)
        check(self._lib.pf_engine_peer_export(self._h, buf), "pf_engine_peer_export")
        return buf.raw

    def peer_attach(self, handles, rank: int, nranks: int) -> None:
        """`handles`: the 64-byte handles of all ranks in rank order; selects the in-kernel exchange."""
        blob = b"".join(bytes(h) for h in handles)
        assert len(blob) == 64 * nranks
        buf = C.create_string_buffer(blob, len(blob))
        check(self._lib.pf_engine_peer_attach(self._h, buf, rank, nranks), "pf_engine_peer_attach")

    def peer_status(self) -> None:
        check(self._lib.pf_engine_peer_status(self._h), "pf_engine_peer_status")

    @property
    def launch_count(self) -> int:
        return int(self._lib.pf_engine_launch_count(self._h))

    @property
    def resident_stats(self):
        """(eligible, calls served by the resident kernel, times it was launched) -- pf_engine_resident_stats."""
        calls, launches = C.c_uint64(0), C.c_uint64(0)
        ok = self._lib.pf_engine_resident_stats(self._h, C.byref(calls), C.byref(launches))
        return bool(ok), int(calls.value), int(launches.value)

    def close(self) -> None:
        if self._h:
            self._lib.pf_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def peer_attach_local(engines) -> None:
    """Attach the exchange buffers of engines that live in THIS process (rank = position in the list)."""
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    check(_lib.load().pf_engine_peer_attach_local(arr, len(engines)), "pf_engine_peer_attach_local")


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(_lib.load().pf_nccl_unique_id(buf), "pf_nccl_unique_id")
    return buf.raw


def blasso_log_prior(theta, tau: float = 5.0) -> float:
    """pyblasso.py:123-133 (host side, O(p))."""
    t = _c(theta)
    return float(_lib.load().pf_blasso_log_prior(t.ctypes.data_as(_dp), t.size - 1, tau))
