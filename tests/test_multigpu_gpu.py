"""Multi-GPU correctness under the driver's `pytest -m gpu`: self-spawns torchrun over the visible GPUs
(2 are enough) and runs tools/multigpu_check.py -- sharded search through the fused NVLink gather and the
NCCL variant against the oracle, short shards, two streams with more queries than SMs, overflow fallback,
sharded self-join.  Skipped on a single-GPU box (logs of 2- and 8-GPU runs are kept under profiles/)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_search_and_selfjoin_on_all_visible_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"{n} GPU visible: the sharded path needs at least 2 (run under `gpurun --gpus 2`)")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tools" / "multigpu_check.py")]
    env = dict(os.environ, MMRS_GATHER_TIMEOUT_MS="60000", OMP_NUM_THREADS="4")
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail
    assert f"multigpu ok: world={world}" in res.stdout, tail
