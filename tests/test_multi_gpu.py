"""GPU, >= 2 devices: NCCL all-to-all transposes and the sharded radial loop (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lm_pipeline_single_rank():
    """The chunk-pipelined LM -> LM calls (device and host containers, MHD and double-curl hydro sets) on ONE rank: the
    transposes degenerate to the lo <-> st permutation of every level chunk, everything else is the multi-rank code path.
    MAGIC_LM_TAPER=2 forces short first / last chunks, which the multi-rank runs use by default."""
    env = dict(os.environ, MAGIC_LM_TAPER="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1", "--master-addr", "127.0.0.1",
           "--master-port", "29532", os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:]


def test_nccl_transpose_and_sharded_loop():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:]
