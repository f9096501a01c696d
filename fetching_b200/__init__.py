"""fetching_b200 -- B200-native frontier log-likelihood engine for prefetching Metropolis-Hastings.

Drop-in for the evaluation path of elaine84/fetching: the C ABI is include/pfetch.h, built into
fetching_b200/lib/libpfetch.so (hand-written sm_100a kernels).  Importing this package never
touches `oracle/`; it raises if the CUDA library is absent when an engine is requested.
"""
from ._lib import (F32, F64, MODEL_BLASSO, MODEL_BOLTZMANN_DENSE, MODEL_BOLTZMANN_SCALAR, MODEL_MIXTURE,
                   MODEL_NORMAL, OPT_ASSUME_IDENTITY_COV, OPT_NODE_SPLIT, OPT_PEER_EXCHANGE, OPT_MATVEC_TENSOR,
                   PfetchError)
from .engine import FrontierEngine, blasso_log_prior, nccl_unique_id, peer_attach_local

__all__ = ["FrontierEngine", "blasso_log_prior", "nccl_unique_id", "PfetchError", "F32", "F64", "MODEL_NORMAL",
           "MODEL_BOLTZMANN_SCALAR", "MODEL_BOLTZMANN_DENSE", "MODEL_MIXTURE", "MODEL_BLASSO",
           "OPT_ASSUME_IDENTITY_COV", "OPT_NODE_SPLIT", "OPT_PEER_EXCHANGE", "OPT_MATVEC_TENSOR", "peer_attach_local"]
