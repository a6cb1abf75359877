"""Multi-GPU parity (skips below 2 GPUs): tools/check_multi.py under torchrun -- banded frames assembled by the library's NCCL
all-gather, torch's all-gather and the fused multicast / peer stores of the fine kernel, each against a single-device render
of the whole frame, bit for bit on every rank."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 8])
def test_banded_frames_on_every_rank(world):
    n = _n_gpus()
    if n < world:
        pytest.skip(f"needs {world} GPUs, found {n}")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", GG_CHECK_PATHS="1500")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "check_multi.py")],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "CHECK_MULTI OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
