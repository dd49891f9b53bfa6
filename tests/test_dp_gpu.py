"""Data parallel on real GPUs (needs >= 2 B200s on one NVSwitch box; skipped otherwise): the in-switch
reduce + Adam kernel against NCCL all_reduce + the flat Adam kernel, run under torchrun
(tools/dp_parity.py: gradients bit-identical sums, parameters / Adam state within fp32 rounding, replicas
identical across ranks)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_switch_reduce_adam_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "dp_parity.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = res.stdout + res.stderr
    if "SKIP: no NVSwitch multicast" in out:
        pytest.skip("no multicast support")
    assert res.returncode == 0 and "DP PARITY OK" in out, out[-3000:]
