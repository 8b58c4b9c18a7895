"""Multi-GPU parity under -m gpu (skipped on boxes with fewer than two GPUs): torchrun, one process per GPU, NCCL.
See tests/mgpu_worker.py for what is compared.  Also the low-host-thread wire format (raw floats + clamp mask),
which is what a rank uses when 8 ranks share a 16-CPU host."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2])
def test_sharded_run_equals_single_gpu_run(torch_cuda, world):
    if torch_cuda.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "MGPU-OK world=%d" % world in r.stdout, r.stdout[-2000:]


def test_host_pass_with_two_host_threads_matches_default(torch_cuda, tmp_path):
    """DSRG_B200_HOST_THREADS=2 (what each of 8 ranks gets on a 16-CPU host): cues and seeds travel as floats, the
    in-place clamp as a bit mask.  Same outputs as the default wire format."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from dsrg_b200 import api, synth
b = synth.make_batch(5, 41, 41, cues="cam", image="smooth", start=20)
b["probs"][1, 3, 4:9, 5] = 1e-6          # values the in-place clamp must lift
eng = api.Engine(5, 41, 41, 21)
p = b["probs"].copy()
q = np.empty_like(p)
seeds = eng.dsrg_forward_host(b["labels"], p, b["cues"], b["image"], api.crf_params(12.0), 0.99, 0.85, crf_out=q)
np.savez(sys.argv[1], seeds=seeds, probs=p, q=q)
''' % ROOT
    outs = []
    for threads in ("2", "16"):
        f = str(tmp_path / ("o%s.npz" % threads))
        env = dict(os.environ, DSRG_B200_HOST_THREADS=threads)
        r = subprocess.run([sys.executable, "-c", code, f], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-3000:]
        outs.append(np.load(f))
    a, b = outs
    assert np.array_equal(a["probs"], b["probs"]) and a["probs"].min() >= np.float32(1e-4)
    assert np.abs(a["q"] - b["q"]).max() <= 2e-5
    assert int((a["seeds"] != b["seeds"]).sum()) <= 2
