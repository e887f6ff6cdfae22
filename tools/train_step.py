#!/usr/bin/env python
"""BASELINE.json config 4: SAST Gen1 training step of the recurrent backbone under DDP.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py [--seq 21]

Per rank: batch 4, L-frame sequence with BPTT through the LSTM states (train.py / modules/detection.py:113-221 protocol),
AdamW, gradient all-reduce by stock DistributedDataParallel over NCCL (the reference's strategy, train.py:91-98).  The
detection head and its SimOTA loss are out of scope, so the loss here is a fixed random projection
of the stage-2..4 features -- every backbone parameter receives a gradient.
The SAST blocks run through torch.ops.sast.score_fwd / select / layer_fwd with their registered hand-written backward
(sast_score_bwd / sast_layer_bwd); --path torch switches to the dense torch-autograd statement for an A/B.
Reported: step time (max over ranks), frames/s, and the gradient all-reduce: its size, its exposed share (step time
with DDP minus step time of the same ranks without gradient synchronisation, `no_sync`) and a stand-alone NCCL
all-reduce of the same number of bytes."""
import argparse
import contextlib
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--seq", type=int, default=21)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--path", default="kernels", choices=["kernels", "torch"])
ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
args = ap.parse_args()
os.environ["SAST_B200_TRAIN"] = args.path

import sast_b200  # noqa: E402
from sast_b200 import _lib as L  # noqa: E402
from sast_b200 import parallel  # noqa: E402
from sast_b200.config import backbone_config  # noqa: E402

rank, world, local = parallel.env_rank()
device = torch.device("cuda", local)
torch.cuda.set_device(device)
parallel.init("nccl", device)
torch.manual_seed(0)
net = sast_b200.build_recurrent_backbone(backbone_config((256, 320), partition_split_32=1)).to(device).train()
for mod in net.modules():
    if isinstance(mod, sast_b200.MS_WSA):
        mod.precision = L.FP32 if args.precision == "fp32" else L.BF16
model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], gradient_as_bucket_view=True) if world > 1 else net
opt = torch.optim.AdamW(net.parameters(), lr=2e-4)
n_params = sum(p.numel() for p in net.parameters())
g = torch.Generator().manual_seed(100 + rank)
frames = [(torch.rand(args.batch, 20, 256, 320, generator=g) > 0.9).to(torch.uint8).to(device) for _ in range(args.seq)]
proj = {s: torch.randn(net.stage_dims[s - 1], device=device) for s in (2, 3, 4)}


def step(sync=True):
    ctx = contextlib.nullcontext() if (sync or world == 1) else model.no_sync()
    with ctx:
        states, loss = None, 0.0
        for x in frames:
            feats, states, _ = model(x, states)
            loss = loss + sum((feats[s] * proj[s].view(1, -1, 1, 1)).mean() for s in (2, 3, 4))
        opt.zero_grad(set_to_none=True)
        loss.backward()
    torch.nn.utils.clip_grad_value_(net.parameters(), 1.0)
    opt.step()
    return float(loss.detach())


def timed(sync):
    for _ in range(args.warmup):
        step(sync)
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step(sync)
    torch.cuda.synchronize()
    dt, = parallel.reduce_scalars([(time.perf_counter() - t0) / args.steps], "max", device)
    return dt, loss


dt, loss = timed(True)
dt_nosync, ar_alone = None, None
if world > 1:
    dt_nosync, _ = timed(False)
    flat = torch.zeros(n_params, device=device)
    for _ in range(3):
        dist.all_reduce(flat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(flat)
    e1.record()
    torch.cuda.synchronize()
    ar_alone, = parallel.reduce_scalars([e0.elapsed_time(e1) / 10 * 1e-3], "max", device)
if rank == 0:
    missing = [n for n, p in net.named_parameters() if p.grad is None]
    print(json.dumps({"workload": "gen1_train_backbone", "path": args.path, "precision": args.precision, "n_gpus": world,
                      "batch_per_gpu": args.batch, "seq": args.seq, "s_per_step": dt, "frames_per_s": world * args.batch * args.seq / dt,
                      "loss": loss, "params": n_params, "grad_allreduce_mb": n_params * 4 / 1e6,
                      "s_per_step_no_grad_sync": dt_nosync,
                      "allreduce_exposed_s": None if dt_nosync is None else max(dt - dt_nosync, 0.0),
                      "allreduce_standalone_s": ar_alone,
                      "allreduce_standalone_busbw_gbs": None if not ar_alone else n_params * 4 * 2 * (world - 1) / world / ar_alone / 1e9,
                      "params_without_grad": missing, "launches_ours": int(L.lib().sast_launch_count())}))
if world > 1:
    dist.destroy_process_group()
