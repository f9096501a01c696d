"""Cross-GPU parity under torchrun (one process per GPU): row shards and node split, the in-kernel NVLink peer
exchange and the NCCL all-reduce / all-gather, against a single engine holding the whole data set.  Needs >= 2 GPUs
(`gpurun --gpus N`); on a one-GPU box the virtual-rank tests of test_gpu_peer_exchange.py cover the protocol and
bench.py checks every rank's shard against the oracle at every N."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_and_node_split_engines_agree_with_one_gpu():
    import torch
    have = torch.cuda.device_count()
    if have < 2:
        pytest.skip("needs >= 2 GPUs (have %d)" % have)
    n = 8 if have >= 8 else 4 if have >= 4 else 2
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
                        os.path.join(ROOT, "tests", "multigpu_check.py")], capture_output=True, text=True, timeout=900)
    print(p.stdout[-4000:])
    assert p.returncode == 0, p.stderr[-4000:]
    assert "MULTI-GPU CHECK OK world %d" % n in p.stdout
