#!/usr/bin/env python
"""bench.py -- distillation-step throughput (face + audio pairs / s) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU arm (oracle restatement of the MatConvNet path)

One step = SENet50-ferplus teacher forward on a batch of 48x48 uint8 faces (preprocessing fused on the
device) -> max-aggregation of the frame logits -> VGGVox student forward + backward on 512x300
spectrograms with the T=2 softmax cross-entropy -> (all-reduce of the 16.6 M-parameter gradient over
NCCL when N > 1) -> SGD-momentum update.

Scaling.  BASELINE.json's headline configuration (C4) is ONE batch of 256 pairs split over the GPUs
(cnn_train_dag gives lab i the samples batch(i:numlabs:end), emoVoxCeleb/run_distillation.m:75,179-181), so the
default is STRONG scaling: `--global-batch 256`, 256 / N pairs per GPU.  The weak-scaling measurement (256 pairs
on every GPU) runs in the same process when N > 1 and is reported under "weak_scaling"; `--scaling weak` makes it
the headline instead.

Other BASELINE configs: `--config c2` (ResNet50 teacher forward, 256 x 224 x 224 x 3), `--config c3` (student
forward + backward + update, batch 128 @512x300), `--config c5` (embedding sweeps, batch 64...1024).
At N = 1 the line also carries "parity_mode": the student step of the same workload in the fp32-equivalent mode
(split-operand convolutions, parity.StudentProgramF32 -- the configuration whose gradients meet 1e-3), beside the
fp16-operand fast mode the headline is measured in.

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


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_seconds(config, units, threads, teacher="senet50"):
    """One pass of `config` through the CPU restatement (oracle port, fp32 torch CPU kernels) on `units` samples."""
    import torch

    from oracle import nets

    torch.set_num_threads(threads)
    if config == "c4":
        tp, sp = nets.teacher_init(teacher), nets.student_init()
        faces = nets.faces48_to_input(nets.synth_faces48(units))
        spec = nets.synth_spectrograms(units, WIDTH)
        t0 = time.perf_counter()
        logits = nets.teacher_forward(tp, faces, nets.TorchOps)
        target = np.stack([nets.aggregate_logits(logits[0, 0, :, i][None, :]) for i in range(units)], axis=1).reshape(1, 1, 8, units)
        nets.distillation_student_step(sp, {}, spec, target.astype(np.float32), lr=1e-4, ops=nets.TorchOps)
    elif config == "c2":
        tp = nets.teacher_init("resnet50")
        faces = nets.synth_faces(units)
        t0 = time.perf_counter()
        nets.teacher_forward(tp, faces, nets.TorchOps)
    elif config == "c3":
        sp = nets.student_init()
        spec, tgt = nets.synth_spectrograms(units, WIDTH), nets.synth_teacher_logits(units)
        t0 = time.perf_counter()
        nets.distillation_student_step(sp, {}, spec, tgt, lr=1e-4, ops=nets.TorchOps)
    else:   # c5: one face through the teacher and one clip through the student in test mode, per unit
        tp, sp = nets.teacher_init("senet50"), nets.student_randomize_bn(nets.student_init())
        faces, spec = nets.synth_faces(units), nets.synth_spectrograms(units, WIDTH)
        t0 = time.perf_counter()
        nets.teacher_forward(tp, faces, nets.TorchOps)
        nets.student_forward(sp, spec, "test", nets.TorchOps)
    return time.perf_counter() - t0


CONFIGS = {
    "c4": dict(metric=METRIC, unit=UNIT,
               workload="full distillation step: %s-ferplus teacher fwd (48x48 uint8 faces -> 224x224x3) + VGGVox student "
                        "fwd+bwd @512x300 + T=2 softmax CE + SGD-momentum"),
    "c2": dict(metric="ResNet50-ferplus teacher forward faces/sec", unit="faces/s",
               workload="ResNet50-ferplus teacher forward (test-mode BN folded), 224x224x3 single faces"),
    "c3": dict(metric="VGGVox student step clips/sec", unit="clips/s",
               workload="VGGVox student forward + backward + T=2 softmax CE + SGD-momentum on 512x300 spectrograms"),
    "c5": dict(metric="embedding extraction samples/sec (SENet50 face logits + VGGVox audio logits)", unit="samples/s",
               workload="compute_visual_feats / compute_audio_feats: SENet50 forward on 224x224x3 faces and VGGVox test-mode "
                        "forward on 512x300 spectrograms, one face + one clip per sample"),
}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (no MATLAB here -> the oracle port), all host threads, rank 0 only."""
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    t_probe = cpu_seconds(args.config, 1, threads, args.teacher)
    budget = 150.0 / max(1, args.steps + args.warmup)
    units = int(max(1, min(8, budget / max(t_probe, 1e-3))))
    for _ in range(args.warmup):
        cpu_seconds(args.config, units, threads, args.teacher)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_seconds(args.config, units, threads, args.teacher)
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    sample = "%d sample(s) per step of: %s; fp32, torch CPU kernels behind the MatConvNet operator semantics" % (
        units, cfg["workload"] % args.teacher if "%s" in cfg["workload"] else cfg["workload"])
    print(json.dumps({
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"] % args.teacher if "%s" in cfg["workload"] else cfg["workload"],
                   "samples_per_step": units, "spectrogram": "512x300", "faces": "48x48 uint8 -> 224x224x3" if args.config == "c4" else "224x224x3"},
        "cpu_baseline": {"value": value, "unit": cfg["unit"], "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm helpers
class Dist:
    """rank / world plumbing: barrier + max-over-ranks timing as the bench contract asks."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.allreduce = (lambda g: dist.all_reduce(g)) if self.world > 1 else None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_loop(D, stream, fn, steps, warmup):
    """`warmup` untimed + `steps` timed calls of fn between barriers; CUDA events on `stream`; max over ranks (ms total)."""
    torch = D.torch
    for _ in range(warmup):
        fn()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    D.barrier()
    return D.max_over_ranks(e0.elapsed_time(e1))


class StepDriver:
    """The timed object: the full distillation step behind the graph-level C ABI (net.DistillStep -> xemo_distill_step, one
    graph replay per step; the gradient exchange is ncclAllReduce issued by the library inside the captured step).  torch
    is used here for pinned host memory, the copy stream and the CUDA events that time the region -- nothing else."""

    def __init__(self, D, args, B, global_batch):
        import torch

        from mcncrossmodalemotions_b200 import zoo
        from mcncrossmodalemotions_b200.net import Comm, DistillStep

        self.torch, self.D, self.B = torch, D, B
        dev = torch.device("cuda", D.local)
        self.stream = torch.cuda.Stream(dev)
        self.copy_stream = torch.cuda.Stream(dev)
        self.step = DistillStep(zoo.teacher_init(args.teacher), zoo.student_init(), B, WIDTH, device=D.local, stream=self.stream.cuda_stream)
        self.ctx = self.step.ctx
        if D.world > 1:
            self.step.comm = Comm.from_torch(self.ctx)
        self.step.student.set_hyper(lr=1e-4, momentum=0.9, weight_decay=5e-4, batch_size=global_batch)
        self.step.student.set_overlap(args.overlap)
        self.faces_h, self.spec_h = synth_inputs(B, D.rank)
        with torch.cuda.stream(self.stream):
            self.stage_faces = torch.empty(self.faces_h.numel(), dtype=torch.uint8, device=dev)
            self.stage_spec = torch.empty(self.spec_h.numel(), dtype=torch.float32, device=dev)
        self.loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.h2d_done, self.stage_free = torch.cuda.Event(), torch.cuda.Event()
        self.stage_free.record(self.stream)
        self.h2d_bytes = self.faces_h.numel() + self.spec_h.numel() * 4
        self.d2h_bytes = 8
        self.scalars = self.step.student.buffer("scalars")

    def prefetch(self):
        """asynchronous H2D of the next step's inputs (pinned host tensors) on the copy stream"""
        torch = self.torch
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.stage_free)
            self.stage_faces.copy_(self.faces_h, non_blocking=True)
            self.stage_spec.copy_(self.spec_h, non_blocking=True)
            self.h2d_done.record(self.copy_stream)

    def step_resident(self):
        self.step.step()

    def step_host(self):
        """consume the prefetched inputs, run the step, copy objective / classerror back to pinned host memory"""
        import ctypes as C

        self.stream.wait_event(self.h2d_done)
        self.step.teacher.set_input(self.stage_faces)
        self.step.student.set_input(self.stage_spec)
        self.stage_free.record(self.stream)
        self.step.step()
        self.ctx.d2h(C.c_void_p(self.loss_host.data_ptr()), C.c_void_p(self.scalars), 8)

    def sync(self):
        self.ctx.sync()

    def close(self):
        # networks first: their captured graphs hold NCCL nodes, and ncclCommDestroy waits for those graphs to be destroyed
        self.step.student.close()
        self.step.teacher.close()
        if self.step.comm:
            self.step.comm.close()


def measure_step(D, args, B, global_batch, clocks=None, with_e2e=True):
    """Resident and end-to-end timing of the distillation step at per-GPU batch B.  Returns (driver, dict)."""
    import torch

    drv = StepDriver(D, args, B, global_batch)
    drv.prefetch()
    drv.step_host()
    drv.sync()
    run = drv.step_resident
    for _ in range(args.warmup):
        run()
    D.barrier()
    t0 = clocks.mark() if clocks else None
    c0 = drv.ctx.launch_count()
    ms_res = timed_loop(D, drv.stream, run, args.steps, 0)
    launches = drv.ctx.launch_count() - c0
    out = dict(ms_res=ms_res, launches=launches)
    if with_e2e:
        # end to end: pinned host buffers -> H2D -> step -> D2H loss, every step (copies of step i+1 overlap step i)
        for _ in range(2):
            drv.prefetch()
            drv.step_host()
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(drv.stream)
        drv.prefetch()
        for i in range(args.steps):
            drv.step_host()
            if i + 1 < args.steps:
                drv.prefetch()
        e1.record(drv.stream)
        D.barrier()
        out["ms_e2e"] = D.max_over_ranks(e0.elapsed_time(e1))
    if clocks:
        t1 = clocks.mark()
        # a short timed region (few steps, many GPUs sharing one nvidia-smi) can end before a 200 ms sample lands inside it:
        # keep the same step running (untimed) until two samples have been taken under this load
        extra = 0
        for _ in range(50):
            need = 1.0 if (clocks.proc and clocks.count_between(t0, t1) < 2) else 0.0
            if D.max_over_ranks(need) == 0.0:   # collective decision: every rank runs the same number of extra steps
                break
            for _ in range(8):
                run()
            extra += 8
            D.barrier()
            t1 = clocks.mark()
        clk = clocks.stop(t0, t1)
        clk["extra_load_steps"] = extra
        out["clocks"] = clk
    drv.sync()
    out["loss"] = float(drv.loss_host[0])
    return drv, out


def graph_ms(D, ctx, stream, record, iters=5):
    """Capture `record()` into a CUDA graph and time its replay (ms per launch)."""
    torch = D.torch
    record()
    ctx.capture_begin()
    record()
    g = ctx.capture_end()
    g.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        g.launch()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    g.destroy()
    return ms


def conv_roofline(D, args, B, ms_step):
    """All tcgen05 convolution launches of one step, timed one by one (instrumented eager pass over the Python-side
    assembly of the same kernel sequence: its per-call hooks are what times each launch), plus the fractions north_star
    names: whole step, teacher forward alone, student step alone (graph replays)."""
    import torch

    from mcncrossmodalemotions_b200 import zoo
    from mcncrossmodalemotions_b200.distill import DistillationStep

    step = DistillationStep(zoo.teacher_init(args.teacher), zoo.student_init(), B, WIDTH, device=D.local)
    faces_h, spec_h = synth_inputs(B, D.rank)
    step.prefetch(faces_h, spec_h)
    step.step_host()
    step.sync()
    prof = ConvProfiler(step.stream)
    step.use_graph = False
    side, step.side, step.student.side_stream = step.side, None, None   # per-op timing needs one stream
    step.grad_step()  # warm
    step.sync()
    step.ctx.profiler = prof
    with torch.cuda.stream(step.stream):
        step.grad_step()
    step.ctx.profiler = None
    fl, t, n_launch, by = prof.summary()
    t_ms = graph_ms(D, step.ctx, step.stream, step.teacher._record)
    s_ms = graph_ms(D, step.ctx, step.stream, lambda: (step.student._record_forward(True), step.student._record_backward(),
                                                         step.student._record_update()))
    step.use_graph = True
    step.side, step.student.side_stream = side, side
    peaks = measured_peaks()
    peak = peaks["tflops_sustained"]
    achieved = fl / t / 1e12
    frac_of = lambda gflop_per, ms: B * gflop_per / ms / peak   # GFLOP / ms == TFLOP/s
    arch = step.teacher.arch
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": None, "traffic_note": "dram__bytes of the conv launches: profiles/r02_conv_traffic.json (ncu capture, not measured in this run)",
            "kernel": "conv_fprop_kernel / conv_wgrad_kernel (tcgen05): all %d convolution launches of one step" % n_launch,
            "launches_per_step": n_launch, "conv_ms_per_step": t * 1e3, "by_op": by,
            "step_frac": frac_of(GFLOP_TEACHER[arch] + GFLOP_STUDENT_FWD_BWD, ms_step),
            "teacher_forward": {"ms": t_ms, "tflops": B * GFLOP_TEACHER[arch] / t_ms, "frac": frac_of(GFLOP_TEACHER[arch], t_ms)},
            "student_step": {"ms": s_ms, "tflops": B * GFLOP_STUDENT_FWD_BWD / s_ms, "frac": frac_of(GFLOP_STUDENT_FWD_BWD, s_ms)},
            "peak_source": peaks["source"] + ", sustained bf16", "operand_dtype": "fp16 x fp16 -> fp32 (TMEM)"}


def cpu_baseline(args, config, units):
    threads = os.cpu_count() or 1
    cpu_seconds(config, 1, threads, args.teacher)  # warm
    dt = cpu_seconds(config, units, threads, args.teacher)
    cfg = CONFIGS[config]
    return {"value": units / dt, "unit": cfg["unit"], "cores": threads, "kind": "port",
            "sample": "1 pass over %d sample(s) of the same workload, fp32 torch CPU kernels, %.1f s" % (units, dt)}


def parity_mode_rate(D, args, B=32, steps=2):
    """The student step of the same workload in the fp32-equivalent mode (split-operand convolutions, fp32 storage)."""
    import torch

    from mcncrossmodalemotions_b200 import zoo
    from mcncrossmodalemotions_b200.parity import StudentProgramF32

    rng = np.random.default_rng(5)
    spec = rng.standard_normal((512, WIDTH, 1, B)).astype(np.float32)
    tgt = (3 * rng.standard_normal((1, 1, 8, B))).astype(np.float32)
    prog = StudentProgramF32(zoo.student_init(), B, WIDTH)
    prog.set_input(spec, tgt)
    prog.grad_step(); prog.update(); prog.ctx.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        prog.grad_step(); prog.update()
    prog.ctx.sync()
    dt = (time.perf_counter() - t0) / steps
    prog.ctx.trim()
    del prog
    torch.cuda.empty_cache()
    return {"value": B / dt, "unit": "clips/s", "dtype": "f32 (fp16 x 3 split operands, fp32 accumulate / storage)", "batch": B,
            "ms_per_step": dt * 1e3, "what": "student forward + backward + update through the vl_nn* boundary operators "
            "(parity.StudentProgramF32); the configuration that meets 1e-3 on gradients, tests/test_gpu_parity.py"}


# ------------------------------------------------------------------------------------------------ configs
def run_c4(args, D):
    world = D.world
    strong = args.scaling == "strong"
    if strong:
        if args.global_batch % world:
            raise SystemExit("--global-batch %d is not divisible by %d ranks" % (args.global_batch, world))
        B = args.global_batch // world
    else:
        B = args.per_gpu_batch
    clocks = ClockSampler(D.local)
    clocks.start()   # nvidia-smi needs a few hundred ms to come up: started ahead of the warm-up, rows are time-stamped
    drv, m = measure_step(D, args, B, B * world, clocks)
    kernels = drv.step.num_kernels()
    h2d, d2h = drv.h2d_bytes, drv.d2h_bytes
    drv.close()
    del drv
    D.torch.cuda.empty_cache()
    roof = conv_roofline(D, args, B, m["ms_res"] / args.steps) if D.rank == 0 else None
    D.torch.cuda.empty_cache()
    D.barrier()
    other = None
    if world > 1 and not args.single_line:
        # the other scaling mode in the same process: weak (256 pairs on every GPU) beside the strong headline, or vice versa
        B2 = args.per_gpu_batch if strong else args.global_batch // world
        if B2 != B and B2 >= 1:
            drv2, m2 = measure_step(D, args, B2, B2 * world, None, with_e2e=False)
            drv2.close()
            other = {"scaling": "weak" if strong else "strong", "value": B2 * world * args.steps / (m2["ms_res"] * 1e-3), "unit": UNIT,
                     "ms_per_step": m2["ms_res"] / args.steps, "per_gpu_batch": B2, "global_batch": B2 * world}
            del drv2
            D.torch.cuda.empty_cache()
    cpu = parity = None
    if D.rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            cpu = cpu_baseline(args, "c4", args.cpu_pairs)
        if not args.no_parity_mode:
            parity = parity_mode_rate(D, args)
    D.barrier()
    if D.rank == 0:
        total = B * world * args.steps
        value = total / (m["ms_res"] * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["ms_res"] / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": CONFIGS["c4"]["workload"] % args.teacher, "global_batch": B * world, "per_gpu_batch": B,
                       "parallelism": "dp%d" % world,
                       "l2": "per-step working set (%.1f GB of activations per GPU) exceeds the 126 MB L2; no flush needed" % (0.045 * B),
                       "gflop_per_pair": GFLOP_PAIR, "achieved_tflops_per_gpu": value / world * GFLOP_PAIR / 1e3},
            "e2e": {"value": total / (m["ms_e2e"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": m["ms_e2e"] / args.steps},
            "gpu_launches": int(m["launches"]), "kernels_per_step": kernels, "loss": m["loss"],
            "clocks": m["clocks"], "roofline": roof, "cpu_baseline": cpu,
        }
        if other:
            line["%s_scaling" % other["scaling"]] = other
        if parity:
            line["parity_mode"] = parity
        print(json.dumps(line))


def run_single_program(args, D):
    """c2 / c3 / c5: one network per rank behind the graph-level C ABI (replicas: no collective except c3's gradient sum)."""
    import ctypes as C

    import torch

    from mcncrossmodalemotions_b200 import _lib, zoo
    from mcncrossmodalemotions_b200.net import Comm, StudentNet, TeacherNet

    cfg = CONFIGS[args.config]
    peaks = measured_peaks()
    peak = peaks["tflops_sustained"]
    clocks = ClockSampler(D.local)
    clocks.start()
    rng = np.random.default_rng(100 + D.rank)
    extra = {}
    stream = torch.cuda.Stream(torch.device("cuda", D.local))
    ctx = _lib.Context(D.local, stream.cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr())
    if args.config == "c2":
        B = args.batch or 256
        net = TeacherNet(zoo.teacher_init("resnet50"), B, ctx=ctx)
        host = torch.from_numpy((rng.uniform(0, 255, B * 3 * 224 * 224) - 110.0).astype(np.float32)).pin_memory()
        out_host = torch.zeros(B, 16).pin_memory()
        net.set_input(host); net.run(); net.sync()
        run = net.run

        def run_e2e():
            net.set_input(host)
            net.run()
            ctx.d2h(vp(out_host), C.c_void_p(net.buffer("logits")), out_host.numel() * 4)
        gflop, h2d, d2h = GFLOP_TEACHER["resnet50"], host.numel() * 4, out_host.numel() * 4
    elif args.config == "c3":
        B = args.batch or 128
        net = StudentNet(zoo.student_init(), B, WIDTH, ctx=ctx)
        net.set_hyper(lr=1e-4, batch_size=B * D.world)
        comm = Comm.from_torch(ctx) if D.world > 1 else None
        _, spec_h = synth_inputs(B, D.rank)
        tgt_h = torch.from_numpy((3 * rng.standard_normal((B, 8))).astype(np.float32)).pin_memory()
        loss_host = torch.zeros(2).pin_memory()
        net.set_input(spec_h)
        net._check(net.lib.xemo_net_set_target(net.handle, vp(tgt_h), None))

        def run():
            net.grad_step(comm)
            net.update()

        def run_e2e():
            net.set_input(spec_h)
            net._check(net.lib.xemo_net_set_target(net.handle, vp(tgt_h), None))
            run()
            ctx.d2h(vp(loss_host), C.c_void_p(net.buffer("scalars")), 8)
        run(); net.sync()
        gflop, h2d, d2h = GFLOP_STUDENT_FWD_BWD, spec_h.numel() * 4 + tgt_h.numel() * 4, 8
    else:   # c5 sweep: per batch size, teacher forward + student test-mode forward
        sweep = {}
        sfwd = lambda n_: n_._check(n_.lib.xemo_student_forward(n_.handle, 0, None))
        for B in (64, 128, 256, 512, 1024):
            tnet = TeacherNet(zoo.teacher_init("senet50"), B, ctx=ctx)
            tnet.run(); tnet.sync()
            t_ms = timed_loop(D, stream, tnet.run, 5, 3) / 5
            tnet.close()
            snet = StudentNet(zoo.student_init(), B, WIDTH, ctx=ctx)
            sfwd(snet); snet.sync()
            s_ms = timed_loop(D, stream, lambda: sfwd(snet), 5, 3) / 5
            snet.close()
            sweep["batch %d" % B] = {"teacher_ms": t_ms, "teacher_faces_per_s": B * D.world / t_ms * 1e3, "teacher_frac": B * GFLOP_TEACHER["senet50"] / t_ms / peak,
                                     "student_ms": s_ms, "student_clips_per_s": B * D.world / s_ms * 1e3, "student_frac": B * 5.662228992 / s_ms / peak,
                                     "samples_per_s": B * D.world / (t_ms + s_ms) * 1e3}
        extra["sweep"] = sweep
        B = args.batch or 256
        tnet = TeacherNet(zoo.teacher_init("senet50"), B, ctx=ctx)
        snet = StudentNet(zoo.student_init(), B, WIDTH, ctx=ctx)
        host_f = torch.from_numpy((rng.uniform(0, 255, B * 3 * 224 * 224) - 110.0).astype(np.float32)).pin_memory()
        _, host_s = synth_inputs(B, D.rank)
        out_host = torch.zeros(2, B, 16).pin_memory()
        tnet.run(); sfwd(snet); ctx.sync()

        def run():
            tnet.run()
            sfwd(snet)

        def run_e2e():
            tnet.set_input(host_f)
            snet.set_input(host_s)
            run()
            ctx.d2h(vp(out_host[0]), C.c_void_p(tnet.buffer("logits")), B * 16 * 4)
            ctx.d2h(vp(out_host[1]), C.c_void_p(snet.buffer("pred32")), B * 16 * 4)
        gflop, h2d, d2h = GFLOP_TEACHER["senet50"] + 5.662228992, (host_f.numel() + host_s.numel()) * 4, out_host.numel() * 4
    for _ in range(args.warmup):
        run()
    D.barrier()
    t0 = clocks.mark()
    c0 = ctx.launch_count()
    ms = timed_loop(D, stream, run, args.steps, 0)
    launches = ctx.launch_count() - c0
    ms_e2e = timed_loop(D, stream, run_e2e, args.steps, 2)
    clk = clocks.stop(t0, clocks.mark())
    cpu = cpu_baseline(args, args.config, min(args.cpu_pairs, 4)) if (D.rank == 0 and D.world == 1 and not args.no_cpu_baseline) else None
    D.barrier()
    if args.config == "c3":
        net.close()           # (the network's graphs before the communicator they captured)
        if comm:
            comm.close()
    if D.rank == 0:
        total = B * D.world * args.steps
        value = total / (ms * 1e-3)
        tfl = B * gflop / (ms / args.steps)
        line = {"metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
                "data": "synthetic",
                "config": {"workload": cfg["workload"], "name": args.config, "per_gpu_batch": B, "global_batch": B * D.world,
                           "parallelism": "replicas x%d" % D.world if args.config != "c3" else "dp%d" % D.world,
                           "l2": "activations of one pass exceed the 126 MB L2; no flush needed"},
                "e2e": {"value": total / (ms_e2e * 1e-3), "unit": cfg["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "clocks": clk,
                "roofline": {"bound": "tensor", "achieved": tfl, "peak": peak, "unit": "TFLOP/s", "frac": tfl / peak, "traffic": None,
                             "kernel": "whole pass (all kernels) against the tensor peak: the quantity north_star names",
                             "peak_source": peaks["source"] + ", sustained bf16"},
                "cpu_baseline": cpu}
        line.update(extra)
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--global-batch", type=int, default=256, help="strong scaling: the batch split over the ranks (BASELINE C4)")
    ap.add_argument("--per-gpu-batch", type=int, default=256, help="weak scaling: pairs per GPU")
    ap.add_argument("--batch", type=int, default=0, help="c2 / c3 / c5: per-GPU batch (default: the BASELINE size)")
    ap.add_argument("--teacher", default="senet50", choices=["senet50", "resnet50"])
    ap.add_argument("--cpu-pairs", type=int, default=4, help="samples in the bounded cpu_baseline pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-mode", action="store_true")
    ap.add_argument("--single-line", action="store_true", help="skip the other scaling mode's measurement when N > 1")
    ap.add_argument("--overlap", type=int, default=-1, choices=[-1, 0, 1],
                    help="forked branches inside the captured step (teacher beside student forward, filter gradients beside the "
                         "data-gradient chain): -1 = the library's rule (on for per-GPU batch <= 64), 0 off, 1 on")
    ap.add_argument("--watchdog", type=int, default=int(os.environ.get("XEMO_BENCH_WATCHDOG", "0")),
                    help="dump every thread's Python stack and exit after this many seconds (debugging hangs)")
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler

        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    D = Dist()
    try:
        if args.config == "c4":
            run_c4(args, D)
        else:
            run_single_program(args, D)
    finally:
        D.close()


if __name__ == "__main__":
    main()
