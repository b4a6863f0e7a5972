"""The chunk-record exchange on real GPUs (csrc/gather.cu): two ranks, each decoding its own frames, records to rank 0 through the
NVLink window (copy-engine push and direct stores) and through NCCL; bench.py's own check compares what arrived on rank 0 with
every rank's payload.  Needs two visible GPUs: skipped on a one-GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["window", "window-direct", "nccl"])
def test_records_of_two_ranks_arrive_on_rank_0(kind):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "4", "--warmup", "3", "--frames", "96",
           "--gather", kind, "--no-e2e", "--no-cpu-baseline"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.strip().split("\n") if l.startswith("{")][-1]
    j = json.loads(line)
    assert j["n_gpus"] == 2
    assert "records of all 2 ranks as gathered on rank 0 == their payloads: yes" in j["parity"], j["parity"]
