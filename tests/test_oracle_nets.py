"""Pins the graph restatements (oracle/nets.py) with the shape constraints the reference itself encodes, and
checks the two oracle back-ends against each other."""
import numpy as np
import pytest

from oracle import nets


def test_pool6_table_matches_reference():
    # emoVoxCeleb/emoVoxZoo.m:258-259 / external/compute_audio_feats.m:45-46
    for width, p in nets.POOL6_TABLE.items():
        assert nets.student_pool6_width(width) == p


def test_parameter_count_matches_survey():
    p = nets.student_init()
    conv = sum(v.size for k, v in p.items() if not k.startswith("bn"))
    assert conv == 16_624_456


def test_time2idx():
    # emoVoxCeleb/getBatchEmoVoxCeleb.m:210-214: floor(max(25 t - 1, 0) / 6) + 1
    assert nets.time2idx(0.0) == 1 and nets.time2idx(0.04) == 1 and nets.time2idx(0.28) == 2 and nets.time2idx(4.0) == 17


def test_spectrogram_row_normalisation():
    s = nets.synth_spectrograms(2, 100)
    assert np.allclose(s.mean(axis=1), 0, atol=1e-5) and np.allclose(s.std(axis=1, ddof=1), 1, atol=1e-4)


def test_student_forward_shapes_and_backends_agree():
    p = nets.student_randomize_bn({k: v.astype(np.float64) for k, v in nets.student_init().items()})
    x = nets.synth_spectrograms(2, 100).astype(np.float64)
    y1, tape = nets.student_forward(p, x, "test", nets.NumpyOps, keep=True)
    y2, _ = nets.student_forward(p, x, "test", nets.TorchOps)
    assert y1.shape == (1, 1, 8, 2) and np.allclose(y1, y2, atol=1e-10)
    assert tape["pool1:x"].shape[:3] == (254, 48, 96) and tape["pool6:win"] == (1, 2)


def test_student_step_backends_agree_and_loss_decreases():
    p = {k: v.astype(np.float64) for k, v in nets.student_init().items()}
    x = nets.synth_spectrograms(3, 100).astype(np.float64)
    t = nets.synth_teacher_logits(3).astype(np.float64)
    r1 = nets.distillation_student_step(dict(p), {}, x, t, ops=nets.NumpyOps, update=False)
    r2 = nets.distillation_student_step(dict(p), {}, x, t, ops=nets.TorchOps, update=False)
    assert np.isclose(r1["objective"], r2["objective"])
    for k in ("fc8f", "fc6f", "conv2f", "conv1f", "bn3m", "bn1b"):
        a, b = r1["grads"][k], r2["grads"][k].reshape(r1["grads"][k].shape)
        assert np.abs(a - b).max() <= 1e-9 * np.abs(a).max(), k
    # a few plain SGD steps on one batch reduce the objective (sanity of signs and of the update rule)
    q, st, first = dict(p), {}, None
    for _ in range(4):
        r = nets.distillation_student_step(q, st, x, t, lr=1e-2, ops=nets.TorchOps)
        first = first if first is not None else r["objective"]
    assert r["objective"] < first


def test_sgd_momentum_rule():
    p = {"w": np.array([1.0, -2.0]), "bn1x": np.array([[0.0, 1.0]])}
    g = {"w": np.array([4.0, 8.0]), "bn1x": np.array([[10.0, 3.0]])}
    st = {}
    nets.sgd_momentum(p, st, g, lr=0.1, batch_size=4, momentum=0.9, weight_decay=0.5)
    m = -(0.5 * np.array([1.0, -2.0]) + np.array([1.0, 2.0]))
    assert np.allclose(st["w"], m) and np.allclose(p["w"], np.array([1.0, -2.0]) + 0.1 * m)
    assert np.allclose(p["bn1x"], [[1.0, 1.2]])  # moving average, rate 0.1
    nets.sgd_momentum(p, st, g, lr=0.1, batch_size=4, momentum=0.9, weight_decay=0.5)
    assert np.allclose(st["w"], 0.9 * m - (0.5 * (np.array([1.0, -2.0]) + 0.1 * m) + np.array([1.0, 2.0])))


@pytest.mark.parametrize("arch", ["resnet50", "senet50"])
def test_teacher_shapes(arch):
    p = nets.teacher_init(arch)
    taps = {}
    y = nets.teacher_forward(p, nets.synth_faces(1), nets.TorchOps, taps)
    assert y.shape == (1, 1, 8, 1)
    assert taps["pool1"].shape == (56, 56, 64, 1) and taps["s5b3"].shape == (7, 7, 2048, 1)
    n_conv = sum(1 for k in p if k.endswith("f") and "se" not in k)
    assert n_conv == 54  # 53 convolutions + the 8-way classifier


def test_aggregate_and_face_preprocessing():
    lg = np.array([[1.0, 5.0, 2.0], [3.0, 0.0, 2.0]])
    assert list(nets.aggregate_logits(lg, "max", 3)) == [3, 5, 2] and list(nets.aggregate_logits(lg, "mean", 2)) == [2, 2.5]
    u8 = nets.synth_faces48(2)
    x = nets.faces48_to_input(u8)
    assert x.shape == (224, 224, 3, 2)
    # corner-aligned bilinear: the four corners reproduce the source corners exactly
    assert np.allclose(x[0, 0, :, 0], u8[0, 0, 0] - nets.AVERAGE_IMAGE) and np.allclose(x[223, 223, :, 1], u8[47, 47, 1] - nets.AVERAGE_IMAGE)
