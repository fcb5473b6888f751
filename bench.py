#!/usr/bin/env python
"""bench.py -- distillation-step throughput (face + audio pairs / s) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU arm (oracle restatement of the MatConvNet path)

One step = SENet50-ferplus teacher forward on a batch of 48x48 uint8 faces (preprocessing fused on the
device) -> max-aggregation of the frame logits -> VGGVox student forward + backward on 512x300
spectrograms with the T=2 softmax cross-entropy -> (all-reduce of the 16.6 M-parameter gradient over
NCCL when N > 1) -> SGD-momentum update.  Weak scaling: every rank processes `--per-gpu-batch` pairs.

`value`  : pairs/s with the step inputs already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same metric through DistillationStep.step_host with pinned HOST buffers: H2D of the uint8
           faces + fp32 spectrograms and D2H of the loss inside the timed region (copies of step i+1
           overlap the compute of step i on a second stream).
`roofline`: all tcgen05 convolution launches of one step (fprop / dgrad / wgrad), timed one by one with
           CUDA events in an instrumented eager pass: algorithmic FLOPs / time vs the measured bf16 peak.
`cpu_baseline`: the oracle port (torch CPU kernels behind the MatConvNet operator semantics) on a bounded
           sample, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "distillation step samples/sec (face+audio pair)"
UNIT = "pairs/s"
WIDTH = 300
# algorithmic work (SURVEY.md section 8d / BASELINE.md section 2): conv/fc MACs x 2
GFLOP_TEACHER = {"senet50": 7.716913152, "resnet50": 7.711883264}
GFLOP_STUDENT_FWD_BWD = 16.633
GFLOP_PAIR = GFLOP_TEACHER["senet50"] + GFLOP_STUDENT_FWD_BWD


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(tflops_burst=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm_gbs=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """Host time stamp: rows that arrive between two marks were sampled while the work enqueued between them ran
        (the caller synchronises the device at both marks)."""
        return time.monotonic()

    def count_between(self, t0, t1):
        return sum(1 for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7)

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for t, r in self.rows:
            if len(r) < 7 or (t0 is not None and not (t0 <= t <= t1)):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


class ConvProfiler:
    """before/after hooks of _lib.Context: CUDA events around every tcgen05 convolution launch."""

    def __init__(self, stream):
        import torch

        self.torch, self.stream, self.records, self._cur = torch, stream, [], None

    @staticmethod
    def _flops(name, a):
        v = [x.value if hasattr(x, "value") else x for x in a]
        if name == "op_conv_fwd":
            n, h, w, cin, kout, r, s, sh, sw, pt, pb, pl, pr = v[1], v[2], v[3], v[4], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14]
        elif name == "op_conv_dgrad":
            n, h, w, cin, kout, r, s, sh, sw, pt, pb, pl, pr = v[1], v[2], v[3], v[4], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14]
        elif name == "op_conv_wgrad":
            n, h, w, cin, kout, r, s, sh, sw, pt, pb, pl, pr = v[1], v[2], v[3], v[4], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]
        else:
            return None
        oh, ow = (h + pt + pb - r) // sh + 1, (w + pl + pr - s) // sw + 1
        fl = 2.0 * n * oh * ow * kout * r * s * cin
        # algorithmic (un-padded) work of the re-formulated stems and of the 8-way head
        if cin == 16 and r == 4 and s == 1:
            fl *= 49.0 / 64.0      # student conv1: 7x7x1 taps inside the 4x1x16 space-to-depth filter
        elif cin == 32 and r == 4 and s == 1:
            fl *= 49.0 / 128.0     # student conv1 in pixel-pair form: block-diagonal 4x1x32 filter, 2 x kout columns
        elif cin == 64 and r == 7 and s == 1:
            fl *= 147.0 / 448.0    # teacher conv1 in pixel-pair form: block-diagonal 7x1x64 filter, 2 x kout columns
        elif cin == 32 and r == 7 and s == 1:
            fl *= 147.0 / 224.0    # teacher conv1: 7x7x3 taps inside the 7x1x32 row-im2col filter
        if cin == 128 and r == 5 and s == 5:
            fl *= 96.0 / 128.0     # student conv2: 96 real input channels inside the 128-channel pitch
        if kout == 16 and cin in (1024, 2048):
            fl *= 0.5              # 8 logits padded to 16 output channels
        return fl

    def before(self, name, args):
        fl = self._flops(name, args)
        if fl is None:
            self._cur = None
            return
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        self._cur = (name, fl, e0, e1)

    def after(self, name, args):
        if self._cur:
            self._cur[3].record(self.stream)
            self.records.append(self._cur)
            self._cur = None

    def summary(self):
        self.torch.cuda.synchronize()
        by = {}
        for name, fl, e0, e1 in self.records:
            d = by.setdefault(name, [0.0, 0.0, 0])
            d[0] += fl; d[1] += e0.elapsed_time(e1) * 1e-3; d[2] += 1
        tot_f = sum(d[0] for d in by.values()); tot_t = sum(d[1] for d in by.values()); launches = sum(d[2] for d in by.values())
        return tot_f, tot_t, launches, {k: dict(tflops=d[0] / d[1] / 1e12, ms=d[1] * 1e3, launches=d[2]) for k, d in by.items()}


def synth_inputs(batch, rank):
    """Seeded synthetic step inputs: uint8 48x48 grey faces and row-normalised N(0,1)
    spectrograms, returned in MatConvNet memory order (column-major) as flat pinned torch tensors."""
    import torch

    rng = np.random.default_rng(1000 + rank)
    faces = rng.integers(0, 256, (batch, 48, 48), dtype=np.uint8)                  # [n][w][h] == H x W x N column-major
    spec = rng.standard_normal((batch, WIDTH, 512), dtype=np.float32)              # [n][w][h] == 512 x W x 1 x N column-major
    mu = spec.mean(axis=1, keepdims=True)
    sd = spec.std(axis=1, ddof=1, keepdims=True)                                   # per frequency row over time, N-1 normalised
    spec = (spec - mu) / sd
    return torch.from_numpy(faces.reshape(-1)).pin_memory(), torch.from_numpy(spec.reshape(-1)).pin_memory()


def cpu_step_seconds(pairs, threads):
    """One distillation step of the CPU restatement (oracle port) on `pairs` face+audio pairs."""
    import torch

    from oracle import nets

    torch.set_num_threads(threads)
    tp, sp = nets.teacher_init("senet50"), nets.student_init()
    faces = nets.faces48_to_input(nets.synth_faces48(pairs))
    spec = nets.synth_spectrograms(pairs, WIDTH)
    t0 = time.perf_counter()
    logits = nets.teacher_forward(tp, faces, nets.TorchOps)
    target = np.stack([nets.aggregate_logits(logits[0, 0, :, i][None, :]) for i in range(pairs)], axis=1).reshape(1, 1, 8, pairs)
    nets.distillation_student_step(sp, {}, spec, target.astype(np.float32), lr=1e-4, ops=nets.TorchOps)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (no MATLAB here -> the oracle port), all host threads."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t_probe = cpu_step_seconds(1, threads)
    budget = 150.0 / max(1, args.steps + args.warmup)
    pairs = int(max(1, min(8, budget / max(t_probe, 1e-3))))
    for _ in range(args.warmup):
        cpu_step_seconds(pairs, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step_seconds(pairs, threads)
    dt = time.perf_counter() - t0
    value = pairs * args.steps / dt
    sample = "%d face+audio pair(s) per step (SENet50 fwd + VGGVox fwd/bwd + T-softmax CE + SGD), fp32, torch CPU kernels" % pairs
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full distillation step (SENet50 teacher fwd + VGGVox student fwd/bwd @512x300 + T=2 softmax CE + SGD)",
                   "pairs_per_step": pairs, "spectrogram": "512x300", "faces": "48x48 uint8 -> 224x224x3"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=256)
    ap.add_argument("--teacher", default="senet50", choices=["senet50", "resnet50"])
    ap.add_argument("--cpu-pairs", type=int, default=4, help="pairs in the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from mcncrossmodalemotions_b200 import zoo
    from mcncrossmodalemotions_b200.distill import DistillationStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.per_gpu_batch
    step = DistillationStep(zoo.teacher_init(args.teacher), zoo.student_init(), B, WIDTH, device=local)
    step.student.set_hyper(lr=1e-4, momentum=0.9, weight_decay=5e-4, batch_size=B * world)
    allreduce = (lambda g: dist.all_reduce(g)) if world > 1 else None
    faces_h, spec_h = synth_inputs(B, rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs resident in HBM, warm-up (captures the graphs)
    clocks = ClockSampler(local)
    clocks.start()   # nvidia-smi needs a few hundred ms to come up: started ahead of the warm-up, rows are time-stamped
    step.prefetch(faces_h, spec_h)
    step.step_host(allreduce)
    step.sync()
    for _ in range(args.warmup):
        step.step_resident(allreduce)
    barrier()
    t_load0 = clocks.mark()
    c0 = step.ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(step.stream)
    for _ in range(args.steps):
        step.step_resident(allreduce)
    e1.record(step.stream)
    barrier()
    launches = step.ctx.launch_count() - c0
    ms_res = max_over_ranks(e0.elapsed_time(e1))
    # ---- end to end: pinned host buffers -> H2D -> step -> D2H loss, every step
    for _ in range(2):
        step.prefetch(faces_h, spec_h)
        step.step_host(allreduce)
    barrier()
    e0.record(step.stream)
    step.prefetch(faces_h, spec_h)
    for i in range(args.steps):
        step.step_host(allreduce)
        if i + 1 < args.steps:
            step.prefetch(faces_h, spec_h)
    e1.record(step.stream)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    t_load1 = clocks.mark()
    # a short timed region (few steps, many GPUs sharing one nvidia-smi) can end before a 200 ms sample lands inside it:
    # keep the same step running (untimed) until two samples have been taken under this load
    extra = 0
    for _ in range(50):
        need = 1.0 if (clocks.proc and clocks.count_between(t_load0, t_load1) < 2) else 0.0
        if max_over_ranks(need) == 0.0:   # collective decision: every rank runs the same number of extra steps
            break
        for _ in range(8):
            step.step_resident(allreduce)
        extra += 8
        barrier()
        t_load1 = clocks.mark()
    clk = clocks.stop(t_load0, t_load1)
    clk["extra_load_steps"] = extra
    loss = float(step.loss_host[0])

    # ---- roofline of the tcgen05 convolution launches (instrumented eager pass, rank 0)
    roof = None
    if rank == 0:
        prof = ConvProfiler(step.stream)
        step.use_graph = False
        side, step.side, step.student.side_stream = step.side, None, None   # per-op timing needs one stream
        step.grad_step()  # warm
        step.sync()
        step.ctx.profiler = prof
        with torch.cuda.stream(step.stream):
            step.grad_step()
        step.ctx.profiler = None
        step.use_graph = True
        step.side, step.student.side_stream = side, side
        fl, t, n_launch, by = prof.summary()
        peaks = measured_peaks()
        achieved = fl / t / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
        if os.path.exists(tpath):   # dram__bytes_read+write of the same launches from the committed ncu capture
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("per_gpu_batch") == B:
                traffic = tj["dram_bytes"]
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops_sustained"], "traffic": traffic, "traffic_unit": "bytes per step (all conv launches)", "kernel": "conv_fprop_kernel / conv_wgrad_kernel (tcgen05)",
                "launches_per_step": n_launch, "conv_ms_per_step": t * 1e3, "by_op": by, "peak_source": peaks["source"] + ", sustained bf16",
                "operand_dtype": "fp16 x fp16 -> fp32 (TMEM)"}
    barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_step_seconds(1, threads)  # warm
        dt = cpu_step_seconds(args.cpu_pairs, threads)
        cpu = {"value": args.cpu_pairs / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "1 step of %d pairs (SENet50 fwd + VGGVox fwd/bwd @512x300 + loss + SGD), fp32 torch CPU kernels, %.1f s" % (args.cpu_pairs, dt)}
    if world > 1:
        dist.barrier()
    if rank == 0:
        total = B * world * args.steps
        value = total / (ms_res * 1e-3)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": "full distillation step: %s-ferplus teacher fwd (48x48 uint8 faces -> 224x224x3) + VGGVox student "
                                   "fwd+bwd @512x300 + T=2 softmax CE + SGD-momentum" % args.teacher,
                       "global_batch": B * world, "per_gpu_batch": B, "parallelism": "dp%d" % world,
                       "l2": "per-step working set (>10 GB of activations) exceeds the 126 MB L2; no flush needed",
                       "gflop_per_pair": GFLOP_PAIR, "achieved_tflops_per_gpu": value / world * GFLOP_PAIR / 1e3},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": step.h2d_bytes, "d2h_bytes_per_step": step.d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "kernels_per_step": step.num_kernels(), "loss": loss,
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
