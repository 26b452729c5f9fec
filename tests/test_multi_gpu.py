"""Multi-GPU z-slab path: N>1 ranks must reproduce the single-GPU result (tests/dist_check.py).
Needs ≥2 GPUs, so it is skipped on a one-GPU box; run it with `gpurun --gpus 2 -- python -m pytest
tests/test_multi_gpu.py -m gpu`.  The host-side slab logic has its own CPU test below."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("min_nz,repl_cells", [("8", "0"), ("2", "0"), ("16", None)])   # distributed coarse levels (two depths) / the default plan
def test_slab_parity_vs_single_gpu(min_nz, repl_cells):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 2 if n < 4 else 4
    env = dict(os.environ, WL_DIST_MIN_NZ=min_nz)
    if repl_cells is not None:
        env["WL_DIST_REPL_CELLS"] = repl_cells   # 0: no level is replicated for being small (the default replicates levels ≤ 128³)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_check.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-6000:])
    sys.stderr.write(p.stderr[-3000:])
    assert p.returncode == 0
