"""Two GPUs, NCCL: the sharded single instance (BASELINE configs[3]) with the CUDA evaluators -- the
tcgen05 dense kernel writing straight into the all-reduce buffer, and the observed-entry kernel -- against the
single-GPU evaluation of the whole instance, and the batched maps partitioned over the ranks.  The ranks are
launched the way bench.py's are (torch.distributed.run, one process per GPU) and run tools/check_sharded_nccl.py.
Skipped with fewer than two devices (the driver's 1-GPU box):
    gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu"""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_instance_and_partitioned_maps_on_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_sharded_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["launches"] > 0
    for k, e in out.items():
        if isinstance(e, dict) and "gC" in e:
            assert e["nll"] < 1e-5 and e["gS"] < 1e-4 and e["gC"] < 1e-4, (k, e)
    assert out["partitioned_maps"]["nll"] < 1e-6
