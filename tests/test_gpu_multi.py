"""N > 1 on real GPUs (needs >= 2 devices; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
2 ranks x B clips == 1 rank x 2B clips (train.py:279-281 semantics), see tests/ddp_worker.py."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_equal_one_rank_with_twice_the_batch():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tests", "ddp_worker.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, cwd=ROOT)
    text = out.stdout.decode()
    assert out.returncode == 0, text[-4000:]
    assert "ddp_worker ok" in text, text[-4000:]
