"""Inter-GPU halo path on real GPUs: 2 (and, when present, 4) NCCL ranks run the strict c24 layout (2,2) cases — the
benchmark's decomposition, multi-substep, 8 tracers — and every rank compares its subdomains with the REFERENCE's final
state (tests/multigpu_step.py).  Skipped on boxes with a single GPU; the same script runs on CPU over gloo."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(nproc, case, extra_env=None, port=29531):
    env = dict(os.environ, **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_step.py"), case]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "0 failures" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 4])
def test_strict_step_over_nccl(nproc):
    """Messages through the C ABI (fv3_halo_exchange_nccl on the library's own ncclComm_t; FV3_NATIVE_NCCL=1)."""
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    _launch(nproc, "c24L2k2n3", {"FV3_NATIVE_NCCL": "1", "FV3_EXPECT_TRANSPORT": "fv3_halo_exchange_nccl"}, port=29530 + nproc)


@pytest.mark.gpu
def test_strict_step_over_torch_distributed_p2p():
    """The same case with the messages posted through torch.distributed.batch_isend_irecv (the default transport)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _launch(2, "c24L2k2n3", {"FV3_EXPECT_TRANSPORT": "torch.distributed"}, port=29537)


def test_strict_step_over_gloo_two_processes():
    if torch.cuda.is_available():
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _launch(2, "c24L2k2n3", {"FV3_MULTI_CPU": "1"}, port=29541)
