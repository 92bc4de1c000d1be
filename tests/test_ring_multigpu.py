"""GPU, needs >= 2 devices (skipped on a single-GPU box): ring attention over REAL ranks — one process per GPU under
torchrun, default transport — against the single-GPU kernel and the dense fp32 oracle, per tensor (tools/ring_check.py)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [2048, 8192])
def test_ring_attention_on_real_ranks(n):
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if ngpu >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29577", str(ROOT / "tools" / "ring_check.py"), str(n)]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "RING_CHECK OK" in proc.stdout
