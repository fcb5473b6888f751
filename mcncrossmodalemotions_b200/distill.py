"""The full distillation step on one GPU: SENet50/ResNet50 teacher forward -> per-clip aggregation of the
frame logits -> VGGVox student forward/backward with the temperature-softmax CE -> SGD-momentum update.

The reference runs these as two stages (teacher logits cached offline by
emoVoxCeleb/fetch_emovoxceleb_imdb.m:54-149, then cnn_train_dag at emoVoxCeleb/run_distillation.m:170-182
with getBatchEmoVoxCeleb.m:133-188 selecting and max-pooling the cached frame logits); BASELINE.json's
headline config fuses them, keeping the coupling operator (`max` / `mean` over the frames of the clip,
first numPredEmotions classes, then softmax(. / T) inside the loss)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .programs import StudentProgram, TeacherProgram, _p


class DistillationStep:
    def __init__(self, teacher_params, student_params, batch, width=300, frames_per_clip=1, aggregator="max", device=0,
                 face_input="u8", face_size=48, use_graph=True, grad_scale=1024.0, temperature=2.0, overlap=False, audio_input="spectrogram",
                 comm_overlap=True, comm_split="fc6"):
        self.N, self.F = batch, frames_per_clip
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(self.device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.ctx = _lib.Context(device, self.stream.cuda_stream)
        self.teacher = TeacherProgram(teacher_params, batch * frames_per_clip, device, self.stream, use_graph=False, ctx=self.ctx,
                                      input_mode=face_input, face_size=face_size)
        self.student = StudentProgram(student_params, batch, width, device, self.stream, use_graph=False, grad_scale=grad_scale,
                                      temperature=temperature, ctx=self.ctx, audio_input=audio_input)
        self.audio_key = "wav" if audio_input == "wav" else "spec"
        self.use_mean = 1 if aggregator == "mean" else 0
        # optional second stream: the teacher forward beside the student forward (they only meet at the loss) and the
        # student's filter gradients beside its data-gradient chain.  Measured on B200 (profiles/): no gain -- every
        # kernel already fills the machine (persistent 1-CTA/SM convolutions, full-occupancy HBM kernels) -- so off by default
        self.side = torch.cuda.Stream(self.device) if overlap else None
        self.student.side_stream = self.side
        with torch.cuda.stream(self.stream):
            self.start = torch.arange(0, batch * frames_per_clip, frames_per_clip, dtype=torch.int32, device=self.device)
            self.end = self.start + frames_per_clip
            # staging buffers the copy stream fills while the previous step computes
            self.stage_faces = torch.empty_like(self.teacher.a["faces"])
            self.stage_spec = torch.empty_like(self.student.a[self.audio_key]).view(-1)
            self.loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.use_graph = use_graph
        self.g_grad = self.g_update = self.g_grad_a = self.g_grad_b = None
        # data-parallel overlap: the backward pass is cut after `comm_split` (fc6): the gradients of fc6..fc8 (the tail of
        # the flat buffer, 82 % of its bytes) are all-reduced on a second stream while conv5..conv1 are differentiated
        self.comm_overlap = comm_overlap
        self.split_layer = [L["name"] for L in self.student.layers].index(comm_split)
        self.split_offset = self.student.segs[comm_split + "f"][0]
        self.comm = torch.cuda.Stream(self.device)
        self.ev_a, self.ev_comm = torch.cuda.Event(), torch.cuda.Event()
        self.h2d_done = torch.cuda.Event()
        self.stage_free = torch.cuda.Event()
        self.stage_free.record(self.stream)
        self.h2d_bytes = self.stage_faces.numel() * self.stage_faces.element_size() + self.stage_spec.numel() * 4
        self.d2h_bytes = 8

    # ---- phases
    def _record_grad(self):
        t, s, ctx = self.teacher, self.student, self.ctx
        side = C.c_void_p(self.side.cuda_stream) if self.side is not None else None
        if side:
            ctx.stream_wait(side, None)   # fork
            ctx.set_stream(side)
        t._record()
        ctx.op_logit_aggregate(_p(t.a["logits"]), t.a["logits"].shape[1], _p(self.start), _p(self.end), self.N, s.K, self.use_mean,
                               _p(s.a["target"]))
        if side:
            ctx.set_stream(None)
        s._record_forward(True)
        if side:
            ctx.stream_wait(None, side)   # join: the loss needs the aggregated teacher logits
        s._record_backward()

    def _record_grad_a(self):
        """teacher forward, coupling, student forward, loss, backward of the late layers (split_layer .. fc8)."""
        t, s, ctx = self.teacher, self.student, self.ctx
        t._record()
        ctx.op_logit_aggregate(_p(t.a["logits"]), t.a["logits"].shape[1], _p(self.start), _p(self.end), self.N, s.K, self.use_mean,
                               _p(s.a["target"]))
        s._record_forward(True)
        s._record_backward(lo=self.split_layer, loss=True)

    def _graph(self, attr, record, reset=False):
        if not self.use_graph:
            record()
            return
        g = getattr(self, attr)
        if g is None:
            record()  # eager warm-up (kernel attributes are set outside the capture)
            if reset:
                self.student.reset_metrics()
            self.ctx.capture_begin()
            record()
            g = self.ctx.capture_end()
            setattr(self, attr, g)
            if not reset:
                return  # the warm-up already executed this phase once on the same inputs
        g.launch()

    def grad_step(self):
        if not self.use_graph:
            self._record_grad()
            return
        if self.g_grad is None:
            self._record_grad()  # eager warm-up (kernel attributes are set outside the capture)
            self.student.reset_metrics()
            self.ctx.capture_begin()
            self._record_grad()
            self.g_grad = self.ctx.capture_end()
        self.g_grad.launch()

    def update(self):
        if not self.use_graph:
            self.student._record_update()
            return
        if self.g_update is None:
            self.ctx.capture_begin()
            self.student._record_update()
            self.g_update = self.ctx.capture_end()
        self.g_update.launch()

    def step_resident(self, allreduce=None):
        """One step on the inputs already resident in HBM (teacher.a['faces'], student.a['spec']).  `allreduce(tensor)`
        sums a gradient slice over the data-parallel ranks (NCCL); with comm_overlap the tail slice (fc6..fc8) travels
        while the early layers are still in their backward pass."""
        if allreduce is None:
            self.grad_step()
        elif not self.comm_overlap:
            self.grad_step()
            with torch.cuda.stream(self.stream):
                allreduce(self.student.grad)
        else:
            g = self.student.grad
            self._graph("g_grad_a", self._record_grad_a, reset=True)
            self.ev_a.record(self.stream)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.ev_a)
                allreduce(g[self.split_offset:])
                self.ev_comm.record(self.comm)
            self._graph("g_grad_b", lambda: self.student._record_backward(lo=0, hi=self.split_layer, loss=False))
            with torch.cuda.stream(self.stream):
                allreduce(g[: self.split_offset])
                self.stream.wait_event(self.ev_comm)
        self.update()

    # ---- end-to-end: host buffers in, loss out
    def prefetch(self, faces_host, spec_host):
        """Asynchronous H2D of the next step's inputs (pinned host tensors) on the copy stream."""
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.stage_free)
            self.stage_faces.copy_(faces_host, non_blocking=True)
            self.stage_spec.copy_(spec_host, non_blocking=True)
            self.h2d_done.record(self.copy_stream)

    def step_host(self, allreduce=None):
        """Consume the prefetched inputs, run the step, copy objective / classerror back to pinned host
        memory.  Returns immediately (asynchronous); call sync() before reading loss_host."""
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.h2d_done)
            self.teacher.a["faces"].copy_(self.stage_faces, non_blocking=True)
            self.student.a[self.audio_key].view(-1).copy_(self.stage_spec, non_blocking=True)
            self.stage_free.record(self.stream)
        self.step_resident(allreduce)
        with torch.cuda.stream(self.stream):
            self.loss_host.copy_(self.student.a["scalars"], non_blocking=True)

    def sync(self):
        self.ctx.sync()

    def num_kernels(self):
        return sum(g.num_kernels for g in (self.g_grad, self.g_grad_a, self.g_grad_b, self.g_update) if g)
