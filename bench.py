#!/usr/bin/env python
"""bench.py -- SAST frames/s on synthetic 20-channel event histograms (BASELINE.json metric).

A *step* is one backbone forward (``benchmark.py:51-64`` protocol: ``forward_backbone(x, None)``,
4 stages of conv-downsample -> SAST block -> conv-LSTM) over one batch of B frames per GPU.
Default workload: 1 Mpx, B=8, 384x640 (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 arm (this repo)
  python bench.py --impl reference ...                           # CPU arm: the unmodified reference (baseline/_ref)
                                                                 # on the host cores; --device cuda: eager on the GPU

N>1: launched by torchrun, one rank per GPU; frames are sharded by batch (B per rank, weak
scaling, no collective on the data path); time = max over ranks.
Prints ONE JSON line (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "1mpx_b8": dict(res=(384, 640), batch=8, split=2, desc="SAST 1 Mpx base backbone, batch 8 at 384x640"),
    "gen1_b1": dict(res=(256, 320), batch=1, split=1, desc="SAST Gen1 base backbone, batch 1 at 240x304 padded to 256x320"),
    "1mpx_stream": dict(res=(384, 640), batch=8, split=2, seq=21,
                        desc="SAST 1 Mpx streaming inference, 8 streams per GPU, 21-step recurrent sequences, LSTM state on device"),
}
METRIC = "SAST frames/s @1Mpx 384x640 (backbone forward, benchmark.py protocol)"


def make_inputs(batch, res, sparsity, n_buffers, seed=1, kind="binary"):
    """benchmark.py:58-60: (rand(B,20,H,W) > sparsity) as integers; uint8 here (the dataset's
    own dtype, data/utils/representations.py) so a host->device copy moves 1 byte per bin.
    kind "poisson": event-count histograms clipped at 10 (count_cutoff of the dataset configs) with a fraction
    1 - sparsity of non-empty bins -- the 'realistic' variant of BASELINE config 2."""
    g = torch.Generator().manual_seed(seed)
    if kind == "poisson":
        import math
        lam = -math.log(max(min(sparsity, 1.0 - 1e-9), 1e-9))           # P(count > 0) = 1 - sparsity
        rate = torch.full((batch, 20, res[0], res[1]), lam)
        return [torch.poisson(rate, generator=g).clamp_(max=10).to(torch.uint8) for _ in range(n_buffers)]
    return [(torch.rand(batch, 20, res[0], res[1], generator=g) > sparsity).to(torch.uint8) for _ in range(n_buffers)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_net(workload, precision, device):
    import sast_b200
    from sast_b200 import _lib as L
    from sast_b200.config import backbone_config
    torch.manual_seed(0)
    net = sast_b200.build_recurrent_backbone(backbone_config(workload["res"], embed_dim=64,
                                                             partition_split_32=workload["split"]))
    net = net.to(device).eval()
    prec = L.FP32 if precision == "fp32" else L.BF16
    for m in net.modules():
        if isinstance(m, (sast_b200.MS_WSA, sast_b200.DWSConvLSTM2d)):
            m.precision = prec
    return net


# ----------------------------------------------------------------------------------------------
# Reference arm: the UNMODIFIED reference from baseline/_ref (staged by baseline/make_ref.py; git-ignored, ships with
# the gpurun snapshot) through its own public API, build_recurrent_backbone(cfg).forward, in the timing loop of
# benchmark.py:33-42.  If baseline/_ref is absent the oracle port (same torch ops in the same order) stands in.
# ----------------------------------------------------------------------------------------------
def oracle_cfg(workload):
    mult = 32 * workload["split"]
    return dict(embed_dim=64, dim_multiplier=[1, 2, 4, 8], num_blocks=[1, 1, 1, 1], patch_size=4,
                in_res_hw=list(workload["res"]), partition_size=[workload["res"][0] // mult, workload["res"][1] // mult])


def load_reference():
    """(DictConfig, build_recurrent_backbone) of the staged reference, or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "models")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from omegaconf import DictConfig                                             # the stand-in staged next to it
    from models.detection.recurrent_backbone import build_recurrent_backbone    # reference code, unmodified
    return DictConfig, build_recurrent_backbone


def reference_backbone(workload, state_dict):
    """The reference's RNNDetector for this workload (config/model/sast_yolox/default.yaml + experiment/*/base.yaml
    as config/modifier.py resolves them), carrying `state_dict` (ours: the state-dict layouts are identical)."""
    DictConfig, build = load_reference()
    mult = 32 * workload["split"]
    part = [workload["res"][0] // mult, workload["res"][1] // mult]
    att = dict(partition_size=part, dim_head=32, attention_bias=True, mlp_activation="gelu", mlp_bias=True, mlp_ratio=4,
               drop_mlp=0, drop_path=0, ls_init_value=1e-5, enable_CB=False, AMP=2e-4, BOUNCE=1e-3)
    cfg = DictConfig(dict(
        name="SASTRNN", input_channels=20, enable_masking=False, partition_split_32=workload["split"], embed_dim=64,
        dim_multiplier=[1, 2, 4, 8], num_blocks=[1, 1, 1, 1], T_max_chrono_init=[4, 8, 16, 32], stem=dict(patch_size=4),
        in_res_hw=list(workload["res"]),
        stage=dict(downsample=dict(type="patch", overlap=True, norm_affine=True), attention=att,
                   lstm=dict(dws_conv=False, dws_conv_only_hidden=True, dws_conv_kernel_size=3, drop_cell_update=0))))
    net = build(cfg).eval()
    net.load_state_dict(state_dict, strict=True)
    return net


def our_state_dict(workload):
    import sast_b200
    from sast_b200.config import backbone_config
    torch.manual_seed(0)
    net = sast_b200.build_recurrent_backbone(backbone_config(workload["res"], embed_dim=64,
                                                             partition_split_32=workload["split"]))
    return {k: v.detach() for k, v in net.state_dict().items()}


def cpu_baseline(workload, sparsity, frames, iters, warm, state_dict=None, kind="binary", device="cpu"):
    """frames/s of the reference path on `frames` frames per iteration (benchmark.py:33-42 loop: warm-up, sync,
    timed iterations, sync).  Returns (fps, seconds per iteration, selected tokens per stage, kind)."""
    if state_dict is None:
        state_dict = our_state_dict(workload)
    x = make_inputs(frames, workload["res"], sparsity, 1, kind=kind)[0].int()      # benchmark.py feeds .int()
    if load_reference() is not None:
        net = reference_backbone(workload, state_dict).to(device)
        x = x.to(device)
        fwd = lambda: net(x, None, None)                                          # noqa: E731
        which = "reference"
    else:
        if device != "cpu":
            raise SystemExit("bench.py --impl reference --device cuda needs baseline/_ref (python baseline/make_ref.py)")
        from oracle import sast_oracle as O
        cfg = oracle_cfg(workload)
        fwd = lambda: O.backbone_forward(x, None, state_dict, cfg)                # noqa: E731
        which = "port"

    def sync():
        if device != "cpu":
            torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(warm):
            fwd()
        sync()
        t0 = time.perf_counter()
        for _ in range(iters):
            counts = fwd()[2]
        sync()
        dt = (time.perf_counter() - t0) / iters
    return frames / dt, dt, [int(c) for c in counts], which


def run_reference(args, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    frames = workload["batch"]
    dev = args.device
    fps, dt, counts, which = cpu_baseline(workload, args.sparsity, frames, max(args.steps, 1), max(args.warmup, 1),
                                          kind=args.input, device=dev)
    what = ("unmodified reference (baseline/_ref) build_recurrent_backbone(cfg).forward" if which == "reference"
            else "oracle port of the reference PyTorch path")
    where = f"torch CPU fp32, {torch.get_num_threads()} threads" if dev == "cpu" else "stock PyTorch eager on the GPU, fp32 (TF32 off)"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "device": dev,
        "config": {"workload": args.workload, "desc": workload["desc"], "batch_per_step": frames, "sparsity": args.sparsity,
                   "selected_tokens_per_stage": counts},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads() if dev == "cpu" else 0, "kind": which,
                         "sample": f"{frames} frames per step (full batch), {what}, {where}, benchmark.py:33-42 loop"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def time_steps(fn, steps, device):
    """CUDA events around exactly `steps` calls on the current stream."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) * 1e-3


def run_ours(args, workload):
    from sast_b200 import _lib as L
    from sast_b200 import parallel
    from sast_b200.runner import GraphedBackbone

    rank, world, local = parallel.env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    numa = parallel.bind_to_gpu_numa(local)          # before the pinned buffers below are allocated
    parallel.init("nccl", device)

    B, res = workload["batch"], workload["res"]
    seq = workload.get("seq", 1)                    # frames per stream and step (1: benchmark.py protocol; 21: streaming)
    net = build_net(workload, args.precision, device)
    n_buf = 4
    raw = make_inputs(B, res, args.sparsity, n_buf, seed=1 + rank, kind=args.input)        # every rank: its own frames
    # Host side of the product API: the event histogram bit-packed (1 bit per bin for the binary benchmark input, 4 bits
    # for counts clipped at 10), packed once OUTSIDE the timed region (a data loader's job); expanded on the device INSIDE it (1 bit: by the stem's
    # producer warps, nothing is unpacked to memory; 4 bits: one fp16 NHWC pass).
    bits = {"auto": 1 if args.input == "binary" else 4, "0": 0, "1": 1, "4": 4}[args.pack]
    import sast_b200
    host = [(sast_b200.pack_events(t, bits) if bits else t).pin_memory() for t in raw]
    dev_in = [t.to(device) for t in host]
    lib = L.lib()

    recurrent = seq > 1

    class EagerRunner:
        """Same surface as GraphedBackbone (static input .x, run on it), launching kernel by kernel."""

        def __init__(self):
            self.x = dev_in[0].clone()
            self.states = None

        def reset_states(self):
            self.states = None

        def __call__(self, x=None):
            if x is not None and x.data_ptr() != self.x.data_ptr():
                self.x.copy_(x, non_blocking=True)
            with torch.no_grad():
                f, s, p = net(self.x, self.states if recurrent else None)
            self.states = s
            self.raw_counts = torch.stack([t.reshape(()) for q in p for t, _ in q.terms])
            return f, s, self.raw_counts

    n_runners = 1 if recurrent else 2       # two static input buffers so that the H2D copy of the next batch overlaps compute
    runners = [EagerRunner() if args.no_graph else GraphedBackbone(net, dev_in[0], recurrent=recurrent)
               for _ in range(n_runners)]
    l1 = lib.sast_launch_count()
    with torch.no_grad():
        net(dev_in[0], None)                       # one eager pass just to count our launches per forward
    launches_per_frame_batch = lib.sast_launch_count() - l1

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize(device)

    # ---- kernel-side throughput: inputs already resident in HBM (rotating over n_buf buffers) ----
    def step_resident(i):
        if recurrent:
            runners[0].reset_states()               # a new 21-frame sequence per stream (RNNStates reset semantics)
        for f in range(seq):
            runners[0](dev_in[(i * seq + f) % n_buf])

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_res = time_steps(step_resident, args.steps, device)
    barrier()

    # ---- end to end: pinned host uint8 -> device (copy stream, double buffered against compute),
    #      forward, per-layer selected-token counts back to the host, every frame batch ----
    n_counts = 8
    counts_host = torch.zeros(n_counts, dtype=torch.int32).pin_memory()     # the selection counters' own dtype: a plain D2H copy
    copy_stream = torch.cuda.Stream(device=device)
    nrun = len(runners)
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    stage_in = [runners[k].x for k in range(2)] if nrun == 2 else [dev_in[0].clone() for _ in range(2)]
    main = torch.cuda.current_stream(device)
    for e in ev_done:
        e.record(main)

    def step_e2e(i):
        if recurrent:
            runners[0].reset_states()
        for f in range(seq):
            j = i * seq + f
            k = j % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_done[k])                       # buffer k is free again
                stage_in[k].copy_(host[j % n_buf], non_blocking=True)
                ev_copied[k].record(copy_stream)
            main.wait_event(ev_copied[k])
            raw = runners[k % nrun](stage_in[k])[2]                       # nrun == 2: runs in place on its own static input
            ev_done[k].record(main)
            counts_host.copy_(raw, non_blocking=True)

    for i in range(3):
        step_e2e(i)
    barrier()
    t_e2e = time_steps(step_e2e, args.steps, device)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    t_res, t_e2e = parallel.reduce_scalars([t_res, t_e2e], "max", device)
    frames = B * seq * world * args.steps
    raw = counts_host.tolist()      # selected tokens per SAST layer (2 per stage), whole batch, last frame batch
    counts = [int(raw[2 * i]) // B + int(raw[2 * i + 1]) // B for i in range(4)]

    if rank == 0:
        from roofline import roofline_block            # measured live, same process
        roof = roofline_block(net, workload, args, device, ms_per_step=t_res / args.steps * 1e3)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
            fps_cpu, dt_cpu, _, which = cpu_baseline(workload, args.sparsity, B, 3, 1, sd, kind=args.input)
            cpu = {"value": fps_cpu, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": which,
                   "sample": f"{B} frames x 3 timed iterations (1 warm-up) of the same workload, "
                             + ("unmodified reference (baseline/_ref)" if which == "reference" else "oracle port of the reference")
                             + f" PyTorch CPU path, fp32, {dt_cpu * 1e3:.0f} ms per iteration"}
        line = {
            "metric": METRIC, "value": frames / t_res, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "desc": workload["desc"], "batch_per_gpu": B, "global_batch": B * world,
                       "frames_per_step_per_gpu": B * seq, "sparsity": args.sparsity,
                       "input": ("(rand > sparsity), benchmark.py:58-60" if args.input == "binary" else "Poisson counts clipped at 10, 1 - sparsity of the bins non-empty")
                                + (f", {bits} bit/bin packed" if bits else ", uint8"),
                       "selected_tokens_per_stage": counts, "cuda_graph": not args.no_graph,
                       "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "l2": f"inputs rotate over {n_buf} buffers; per-step working set (activations + workspaces) exceeds the 126 MB L2"},
            "e2e": {"value": frames / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": host[0].numel() * seq * world,
                    "d2h_bytes_per_step": 4 * n_counts * seq * world, "ms_per_step": t_e2e / args.steps * 1e3,
                    "pipeline": "H2D on a copy stream, double buffered against the compute stream",
                    "host_format": (f"event histogram bit-packed {bits} bit/bin (sast_b200.pack_events, outside the timed region); "
                                    "expanded on the device inside it (1 bit: in the stem kernel itself)" if bits else "uint8, 1 byte/bin"),
                    "d2h": "per-layer selected-token counts (the backbone protocol's only host-visible result, benchmark.py:33-42)",
                    "numa": numa},
            "gpu_launches": int(launches_per_frame_batch) * seq * args.steps,
            "gpu_launches_per_frame_batch": int(launches_per_frame_batch),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu (the reference arm) or cuda (the same unmodified code, stock PyTorch eager on the GPU)")
    ap.add_argument("--workload", default="1mpx_b8", choices=sorted(WORKLOADS))
    ap.add_argument("--sparsity", type=float, default=0.0, help="benchmark.py default 0.0 (every pixel active)")
    ap.add_argument("--input", default="binary", choices=["binary", "poisson"],
                    help="binary: (rand > sparsity) as benchmark.py does; poisson: event counts clipped at 10, 1 - sparsity non-empty")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--pack", default="auto", choices=["auto", "0", "1", "4"],
                    help="bits per bin of the host / device input (0: plain uint8; auto: 1 for --input binary, 4 for poisson)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    workload = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, workload)
    else:
        run_ours(args, workload)


if __name__ == "__main__":
    main()
