"""Host-side batch assembly for the student: the teacher -> student coupling and the input tensor layout of
emoVoxCeleb/getBatchEmoVoxCeleb.m (window selection :133-159, time2idx :210-214, aggregation :179-185,
row normalisation :164-169, input list :28-44).  Audio decoding and the spectrogram itself (`runSpec`, VGGVox) are
outside the hot path: callers pass waveform-derived spectrograms, or a `spectrogram_fn`."""
from __future__ import annotations

import numpy as np

FPS, FRAME_STRIDE = 25, 6          # teacher logits exist for every 6th frame of 25 fps video
DATASET_LIMIT_S = 19.9             # getBatchEmoVoxCeleb.m:83-89


def time2idx(t):
    """1-based index into the per-wav logits array of the frame at time `t` seconds (getBatchEmoVoxCeleb.m:210-214)."""
    return int(np.floor(max(t * FPS - 1.0, 0.0) / FRAME_STRIDE) + 1)


def audio_crop_seconds(width, tw_ms=25):
    """audTime = 0.01 * W + 0.001 * Tw - 0.001 (getBatchEmoVoxCeleb.m:67-68): the audio needed for W spectrogram columns."""
    return 0.01 * width + 0.001 * tw_ms - 0.001


def frame_window(num_frames, start_time, end_time):
    """Rows of the F x 8 logits array that fall inside the audio crop [start_time, end_time] (seconds), as a
    0-based half-open range; the end is clamped to the frames that exist (getBatchEmoVoxCeleb.m:143-157)."""
    s = time2idx(start_time)
    e = min(time2idx(end_time), int(num_frames))
    if e < s:
        raise ValueError("audio crop [%g, %g] s selects no teacher frames out of %d" % (start_time, end_time, num_frames))
    return s - 1, e


def aggregate(lgts, method="max", num_pred=8):
    """'max' (run_distillation.m:80) or 'mean' over the frames, first numPredEmotions classes (:179-188)."""
    lgts = np.asarray(lgts, np.float32)
    if method == "max":
        out = lgts.max(axis=0)
    elif method == "mean":
        out = lgts.mean(axis=0)
    else:
        raise ValueError("unrecognised aggregator %s" % method)
    if np.any(np.isnan(out)):
        raise FloatingPointError("NaN teacher logits")  # the reference drops into the debugger here (:189-192)
    return out[:num_pred]


def normalize_rows(spec):
    """inputnorm (getBatchEmoVoxCeleb.m:164-169): per frequency row, (x - mean) / std over time, std with MATLAB's
    default N-1 normalisation.  spec: 512 x W."""
    spec = np.asarray(spec, np.float64)
    mu = spec.mean(axis=1, keepdims=True)
    sd = spec.std(axis=1, ddof=1, keepdims=True)
    return ((spec - mu) / sd).astype(np.float32)


def random_crop_offset(total_samples, crop_samples, rng):
    """Uniform random crop start (1-based sample index, as `randi(wd)`), or 1 with zero padding when the clip is
    shorter than the crop (getBatchEmoVoxCeleb.m:108-119)."""
    wd = int(total_samples) - int(crop_samples)
    return int(rng.integers(1, wd + 1)) if wd >= 1 else 1


def get_batch(spectrograms, wav_logits, crop_times, loss_type="hot-cross-ent", aggregator="max", num_pred=8, inputnorm=True):
    """The `inputs` cell of getBatchEmoVoxCeleb (:28-44) as a dict.
    spectrograms: list of 512 x W arrays (already cropped to the clip's window); wav_logits: list of F_i x 8
    arrays; crop_times: list of (start_s, end_s) of each crop.  Returns data 512 x W x 1 x N, logitTarget
    1 x 1 x num_pred x N, maxLabel 1 x 1 x 1 x N (1-based arg-max of the aggregated logits)."""
    n = len(spectrograms)
    ims, lgs = [], []
    for spec, lg, (t0, t1) in zip(spectrograms, wav_logits, crop_times):
        a, b = frame_window(len(lg), t0, t1)
        lgs.append(aggregate(np.asarray(lg)[a:b], aggregator, num_pred))
        ims.append(normalize_rows(spec) if inputnorm else np.asarray(spec, np.float32))
    data = np.stack(ims, axis=2)[:, :, None, :]
    lgo = np.stack(lgs, axis=1).reshape(1, 1, num_pred, n).astype(np.float32)
    max_label = lgo.argmax(axis=2).reshape(1, 1, 1, n) + 1
    inputs = {"data": data}
    if loss_type == "softmaxlog":
        inputs["maxLabel"] = max_label
    elif loss_type == "euclidean":
        inputs.update(logitTarget=lgo, instanceWeights=np.ones((1, 1, 1, n), np.float32), maxLabel=max_label)
    elif loss_type == "hot-cross-ent":
        inputs.update(logitTarget=lgo, maxLabel=max_label)
    else:
        raise ValueError("unrecognised loss type: %s" % loss_type)
    return inputs


def width_bucket(num_columns, buckets=(100, 200, 300, 400, 500, 600, 700, 800, 900, 1000)):
    """external/compute_audio_feats.m:160-185: the largest bucket not longer than the clip (centre crop)."""
    ok = [b for b in buckets if b <= num_columns]
    if not ok:
        raise ValueError("clip with %d spectrogram columns is shorter than the smallest bucket" % num_columns)
    return max(ok)


def centre_crop(spec, width):
    """external/compute_audio_feats.m:183-186: 1-based rstart = round((W - rsize) / 2) (MATLAB round: halves away from
    zero), 0 mapped to 1; columns rstart : rstart + rsize - 1."""
    w = spec.shape[1]
    rstart = max(int(np.floor((w - width) / 2.0 + 0.5)), 1)
    return spec[:, rstart - 1 : rstart - 1 + width]
