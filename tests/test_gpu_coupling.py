"""Teacher -> student coupling on the device: per-clip window of the cached frame logits
(emoVoxCeleb/getBatchEmoVoxCeleb.m:133-159, time2idx :210-214, end clamped to the frames that exist :151), `max` / `mean`
over the window (:179-188), first numPredEmotions classes (:30), arg-max label (:32)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("method", ["max", "mean"])
@pytest.mark.parametrize("num_pred", [8, 5])
def test_logit_aggregate_ragged_windows(method, num_pred):
    """13-frame windows (3 s crops at 25 fps / stride 6), ragged and clamped at the end of short clips."""
    from oracle import nets
    from mcncrossmodalemotions_b200 import _lib, batch

    rng = np.random.default_rng(21)
    n = 37
    frames = rng.integers(14, 80, n)                     # frames of cached logits per clip (F_i x 8)
    logits = [3.0 * rng.standard_normal((f, 8)).astype(np.float32) for f in frames]
    crop = batch.audio_crop_seconds(300)
    start, end, lens = [], [], []
    base = 0
    for i, f in enumerate(frames):
        # crop start such that some windows run past the last cached frame (the reference clamps endIdx)
        t0 = float(rng.uniform(0, (f * 6 + 1) / 25.0 - 0.5 * crop))
        a, b = batch.frame_window(f, t0, t0 + crop)
        assert 1 <= b - a <= 14
        lens.append(b - a)
        start.append(base + a); end.append(base + b)
        base += f
    assert max(lens) == 13 or max(lens) == 14, lens
    assert min(lens) < 12, "the clamped case must be present"
    flat = np.concatenate(logits, axis=0)
    ref = np.stack([nets.aggregate_logits(flat[a:b], method, num_pred) for a, b in zip(start, end)])
    ctx = _lib.Context(0)
    d_log = torch.from_numpy(flat).cuda()
    d_s, d_e = torch.tensor(start, dtype=torch.int32).cuda(), torch.tensor(end, dtype=torch.int32).cuda()
    out = torch.zeros(n, num_pred, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.op_logit_aggregate(_p(d_log), 8, _p(d_s), _p(d_e), n, num_pred, 1 if method == "mean" else 0, _p(out))
    ctx.sync()
    got = out.cpu().numpy()
    if method == "max":
        assert np.array_equal(got, ref)
    else:
        assert rel_err(got, ref) < 1e-6
    assert np.array_equal(got.argmax(axis=1), ref.argmax(axis=1))      # maxLabel


@pytest.mark.parametrize("method", ["max", "mean"])
def test_distillation_step_with_13_frames_per_clip(method):
    """The fused step with F = 13 teacher frames per clip: the student's target equals the aggregation of the teacher
    program's own frame logits (bit-exact for max) and the oracle's aggregated logits within 1e-3."""
    from oracle import nets
    from mcncrossmodalemotions_b200.distill import DistillationStep

    n, F, width = 4, 13, 100     # (train-mode BN needs more than two samples per batch to be well conditioned)
    tp, sp = nets.teacher_init("senet50"), nets.student_init()
    faces = nets.synth_faces48(n * F)
    spec = nets.synth_spectrograms(n, width)
    step = DistillationStep(tp, sp, n, width, frames_per_clip=F, aggregator=method, use_graph=False)
    step.teacher.set_input(faces)
    step.student.set_input(spec)
    step.student.set_hyper(lr=1e-4, batch_size=n)
    step.step_resident()
    step.sync()
    frame_logits = step.teacher.a["logits"][:, :8].cpu().numpy()
    target = step.student.a["target"].cpu().numpy()
    own = np.stack([nets.aggregate_logits(frame_logits[i * F:(i + 1) * F], method) for i in range(n)])
    if method == "max":
        assert np.array_equal(target, own)
    else:
        assert rel_err(target, own) < 1e-6
    ref_frames = nets.teacher_forward(tp, nets.faces48_to_input(faces), nets.TorchOps).reshape(8, n * F).T
    ref = np.stack([nets.aggregate_logits(ref_frames[i * F:(i + 1) * F], method) for i in range(n)])
    assert rel_err(target, ref) < 1e-3
    m = step.student.metrics()
    out = nets.distillation_student_step(nets.student_init(), {}, spec, ref.T.reshape(1, 1, 8, n).astype(np.float32), ops=nets.TorchOps)
    assert abs(m["objective"] - out["objective"]) <= 1e-3 * abs(out["objective"])
