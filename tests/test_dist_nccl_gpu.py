"""Runs tests/dist_parity_nccl.py under torchrun when the box has >= 2 GPUs (sharded vs unsharded CUDA path)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (run with gpurun --gpus 2)")
def test_nccl_sharded_equals_unsharded():
    n = 2 if torch.cuda.device_count() < 8 else 8
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_parity_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    if r.returncode != 0:       # keep the evidence where a gpurun call brings it back
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "dist_parity_nccl_failure.log"), "w") as f:
            f.write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    report = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0, (report[-1] if report else "") + r.stderr[-1500:]
    assert '"all_ranks": true' in r.stdout
