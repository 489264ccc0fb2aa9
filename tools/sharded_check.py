#!/usr/bin/env python
"""Launched by torchrun with 2+ ranks (tests/test_gpu_multi.py, gpurun --gpus 2): the batch-sharded forward
with the NCCL all-gather must equal the unsharded forward on one GPU bit for bit (eval-mode
outputs depend on the cloud only).  Rank 0 prints one JSON line."""
import contextlib
import io
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv  # noqa: E402
from svnet_b200.parallel import ShardedInference  # noqa: E402
from svnet_b200.synthetic import make_args, one_hot_labels, synthetic_clouds, synthetic_state_dict  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {"world": world}
    cases = [("cls_bin", "SV_DGCNN_CLS", dict(k=20, binary=True), 40, 8, 1024, False),
             ("cls_fp", "SV_DGCNN_CLS", dict(k=20, binary=False), 15, 6, 512, False),
             ("pseg_bin", "SV_DGCNN_PSEG", dict(k=40, binary=True), 50, 4, 2048, True),
             ("cls_bin_tail_batch", "SV_DGCNN_CLS", dict(k=8, binary=True), 40, 1, 128, False)]
    for name, kind, margs, ncls, B, N, lab in cases:
        with contextlib.redirect_stdout(io.StringIO()):
            net = getattr(sv, kind)(make_args(**margs), ncls)
        net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=21))
        net = net.to(dev).eval()
        x = synthetic_clouds(B, N, 22).to(dev)
        ins = (x, one_hot_labels(B).to(dev)) if lab else (x,)
        with torch.no_grad():
            y_all = ShardedInference(net)(*ins)
            y_one = net(*ins)
        ok = torch.equal(y_all, y_one)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        res[name] = {"bit_identical_on_every_rank": bool(flag.item()), "shape": list(y_all.shape)}
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res))
        if not all(v["bit_identical_on_every_rank"] for k, v in res.items() if isinstance(v, dict)):
            sys.exit(1)


if __name__ == "__main__":
    main()
