"""Embedding / logit extraction sweeps: external/compute_visual_feats.m (teacher, batch 128 loop :83-98) and
external/compute_audio_feats.m (student, one clip per eval with the width-bucketed pool6, :116-136, :160-185) --
here the student clips are grouped by bucket and each bucket is batched.  Both run through the graph-level C ABI (net.py)."""
from __future__ import annotations

import numpy as np

from . import batch as B
from .net import StudentNet, TeacherNet


def compute_visual_feats(teacher_params, faces, batch_size=128, device=0, input_mode=None):
    """faces: 224 x 224 x 3 x M normalised singles, or S x S x M uint8 grey faces.  Returns M x 8 logits
    (`gather(squeeze(dag.vars(end).value))'`, compute_visual_feats.m:91-92)."""
    faces = np.asarray(faces)
    u8 = faces.ndim == 3
    m = faces.shape[-1]
    prog = TeacherNet(teacher_params, batch_size, device=device, input_mode="u8" if u8 else "hwcn224",
                      face_size=faces.shape[0] if u8 else 48)
    out = np.zeros((m, prog.K), np.float32)
    for s in range(0, m, batch_size):
        chunk = faces[..., s : s + batch_size]
        n = chunk.shape[-1]
        if n < batch_size:  # last, partial batch: pad (the extra rows are discarded)
            pad = np.zeros(chunk.shape[:-1] + (batch_size - n,), chunk.dtype)
            chunk = np.concatenate([chunk, pad], axis=-1)
        out[s : s + n] = prog.forward(chunk)[:n]
    return out


def compute_audio_feats(student_params, spectrograms, batch_size=64, device=0, inputnorm=True):
    """spectrograms: list of 512 x W_i arrays (whole clips).  Each clip is row-normalised, centre-cropped to the
    largest width bucket <= W_i (compute_audio_feats.m:160-185) and evaluated in test mode with the bucket's pool6
    window (:121-125).  Returns a len(spectrograms) x 8 array."""
    out = np.zeros((len(spectrograms), 8), np.float32)
    buckets = {}
    for i, s in enumerate(spectrograms):
        buckets.setdefault(B.width_bucket(s.shape[1]), []).append(i)
    for width, idxs in sorted(buckets.items()):
        prog = StudentNet(student_params, batch_size, width, device=device)
        for s in range(0, len(idxs), batch_size):
            sel = idxs[s : s + batch_size]
            data = np.zeros((512, width, 1, batch_size), np.float32)
            for j, i in enumerate(sel):
                spec = B.normalize_rows(spectrograms[i]) if inputnorm else np.asarray(spectrograms[i], np.float32)
                data[:, :, 0, j] = B.centre_crop(spec, width)
            out[sel] = prog.forward(data, "test")[: len(sel)]
    return out
