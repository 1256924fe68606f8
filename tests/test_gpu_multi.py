"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one process per GPU under torch.distributed.run,
the product's NCCL swap path, amplitudes assembled across ranks vs the oracle (tests/gpu_multirank_worker.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_circuits_vs_oracle(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "gpu_multirank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_launcher_mode_reproduces_the_golden_dump(world):
    """`hyquas_main file.qasm` started with NO launcher environment drives `world` GPUs by re-executing itself once per GPU
    (host/utils.cpp, the reference's single-process multi-GPU mode, src/utils.cpp:17-60): rank 0's stdout dump must be the
    reference's golden text byte for byte, with nothing (NCCL banners, other ranks' Logger lines) mixed into it."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import re
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "HQ_SPAWNED")}
    env["HQ_NUM_GPUS"] = str(world)
    exe = os.path.join(ROOT, "hyquas_b200", "hyquas_main")
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "qft_28.qasm")], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-800:]
    dump = "".join(l + "\n" for l in r.stdout.splitlines() if not l.startswith("Logger"))
    assert dump == open(os.path.join(ROOT, "tests", "golden", "qft_28.log")).read()
    assert len(re.findall(r"Logger\[\d+\]: Time Cost: \d+ us", r.stdout)) == world
