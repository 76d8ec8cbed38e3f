"""SURVEY §8(e) correctness gate: the landmark-sharded BA over G GPUs equals the single-GPU BA within the stated FP
tolerance (poses 1e-5, points 2e-4, accepted-chi2 trace 1e-6 relative) and its poses are bit-identical on every rank.
Needs >= 2 GPUs on the box (skipped on the single-GPU test tier); one process per GPU through torch.distributed.run."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "sharded_ba_check.py")] + extra
    env = dict(os.environ, NRSLAM_B200_XTIMEOUT_MS="15000")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(lines[-1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("extra", [["--config", "c1"], ["--config", "c3", "--landmarks", "1500", "--keyframes", "10", "--visible", "6"]])
def test_sharded_ba_matches_single_gpu(extra):
    world = min(torch.cuda.device_count(), 8)
    for w in sorted({2, world}):
        out = _run(w, extra)
        assert out["ok"] and out["poses_identical_on_all_ranks"], out
