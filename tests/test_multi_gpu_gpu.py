"""The multi-process slab check (tests/multi_gpu_check.py: one process per GPU, peer-memory and NCCL halo exchange, explicit and
overlapped with the step, every field against the single-GPU run bit for bit) as a pytest: runs wherever `pytest -m gpu` sees at
least two GPUs, skips on a single-GPU box (bench.py --gpus N carries the same comparison in its `parity_check` key)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_multi_process_slabs_match_one_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (one process per GPU)")
    ranks = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    tail = (out.stdout + out.stderr)[-3000:]
    assert out.returncode == 0, tail
    assert f"multi_gpu_check OK on {ranks} ranks" in out.stdout, tail
    assert "mismatches=" in out.stdout and all(line.rstrip().endswith("mismatches=0")
                                               for line in out.stdout.splitlines() if "mismatches=" in line), tail
