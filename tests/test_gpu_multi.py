"""Multi-GPU tests (skipped below 2 visible GPUs; run with `gpurun --gpus 2`): the reference's own
multi-GPU mechanism nn.DataParallel (main_cls_dgcnn.py:125) over two devices, and the batch-sharded
one-process-per-GPU path with the NCCL all-gather against the unsharded forward."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import svnet_oracle as orc
from tests.util import assert_close, quiet, t2n
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")]


def test_data_parallel_two_devices_matches_oracle_and_packs_once():
    import svnet_b200 as sv
    from svnet_b200.sv_layers import _Cached
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=8, binary=True), 40)
    sd = synthetic_state_dict(net.state_dict(), seed=5)
    net.load_state_dict(sd)
    model = torch.nn.DataParallel(net.to("cuda:0"), device_ids=[0, 1]).eval()
    x = synthetic_clouds(4, 128, 5).to("cuda:0")
    with torch.no_grad():
        y = model(x)
        builds = _Cached.BUILDS
        y2 = model(x)                     # fresh replicas, fresh parameter copies: must hit the source-keyed cache
    assert _Cached.BUILDS == builds, "weights were re-packed on the second DataParallel forward"
    assert torch.equal(y, y2)
    with torch.no_grad():
        y_single = net(x)
    assert torch.equal(y, y_single)       # each cloud's logits do not depend on the device or its batch-mates
    rec = {}
    ref = orc.sv_dgcnn_cls(sd, t2n(x), 8, rec=rec)
    with torch.no_grad():
        yf = net(x, forced_idx=[torch.from_numpy(i).to(torch.int32).cuda() for i in rec["idx"]])
    assert_close(t2n(yf), ref, what="forward vs oracle (kNN graphs forced)")


def test_sharded_equals_unsharded_bit_for_bit_nccl():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tools", "sharded_check.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["world"] == 2
    assert all(v["bit_identical_on_every_rank"] for v in res.values() if isinstance(v, dict)), res
