"""The callers either side of the hot path, end to end on the GPU: the run_distillation driver (epochs, lr schedule,
checkpoint + 'continue'), the embedding-extraction sweeps (compute_visual_feats / compute_audio_feats) and the zoo
objects' eval, each against the oracle."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nets():
    from oracle import nets

    return nets


def test_compute_visual_feats_matches_oracle_and_handles_partial_batches(nets):
    from mcncrossmodalemotions_b200 import features, zoo

    u8 = nets.synth_faces48(11)
    p = zoo.teacher_init("senet50-ferplus")
    got = features.compute_visual_feats(p, u8, batch_size=4)            # 4 + 4 + 3 (padded) faces
    ref = nets.teacher_forward(p, nets.faces48_to_input(u8), nets.TorchOps).reshape(8, 11).T
    assert got.shape == (11, 8) and rel_err(got, ref) < 1e-3


def test_compute_audio_feats_buckets_and_batches(nets):
    from mcncrossmodalemotions_b200 import batch as B
    from mcncrossmodalemotions_b200 import features, zoo

    p = nets.student_randomize_bn(zoo.student_init())
    rng = np.random.default_rng(3)
    clips = [rng.standard_normal((512, w)) * 2 + 0.3 for w in (130, 310, 100, 399, 250)]   # buckets 100, 300, 100, 300, 200
    got = features.compute_audio_feats(p, clips, batch_size=2)
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    for i, c in enumerate(clips):
        w = B.width_bucket(c.shape[1])
        x = B.centre_crop(B.normalize_rows(c), w).astype(np.float64)[:, :, None, None]
        ref, _ = nets.student_forward(p64, x, "test", nets.TorchOps)
        assert rel_err(got[i], ref.reshape(8)) < 1e-3, i


def test_zoo_models_eval(nets):
    from mcncrossmodalemotions_b200 import zoo

    teacher = zoo.ferPlusZoo("resnet50-ferplus")
    teacher.move("gpu")
    x = nets.synth_faces(3)
    ref = nets.teacher_forward(teacher.params, x, nets.TorchOps).reshape(8, 3).T
    assert rel_err(teacher.eval({"data": x}), ref) < 1e-3
    student = zoo.emoVoxZoo("emovoxceleb-student", scratch=True, numSeconds=3)
    assert student.pool6 == (1, 8) and student.meta["normalization"]["imageSize"] == (512, 300, 1)
    with pytest.raises(RuntimeError):
        student.move("cpu")
    with pytest.raises(ValueError):
        zoo.emoVoxZoo("no-such-model")


@pytest.mark.parametrize("deterministic", [False, True])
def test_run_distillation_trains_checkpoints_and_resumes(nets, tmp_path, monkeypatch, deterministic):
    """deterministic = True (XEMO_DETERMINISTIC=1: no split-K in the filter gradients, ordered loss / bias sums): a resumed
    run reproduces the uninterrupted one BIT FOR BIT; the default mode differs by the summation order of its atomics."""
    from mcncrossmodalemotions_b200 import batch as B
    from mcncrossmodalemotions_b200 import train as T

    monkeypatch.setenv("XEMO_DETERMINISTIC", "1" if deterministic else "0")
    rng = np.random.default_rng(0)
    n_wavs = 24
    imdb = {"spec": [rng.standard_normal((512, 100)) for _ in range(n_wavs)],
            "wavLogits": [3 * rng.standard_normal((rng.integers(5, 9), 8)).astype(np.float32) for _ in range(n_wavs)]}

    def get_batch(imdb, idx):
        return B.get_batch([imdb["spec"][i] for i in idx], [imdb["wavLogits"][i] for i in idx], [(0.0, 1.0)] * len(idx))

    common = dict(numSeconds=1, batchSize=8, train=np.arange(16), val=np.arange(16, 24), miniVal=1.0, miniEpochRatio=1.0,
                  learningRate=np.full(3, 1e-3))
    p2, info2 = T.run_distillation(imdb, get_batch, root=str(tmp_path), numEpochs=2, **common)
    assert len(info2["train"]) == 2 and info2["train"][1]["objective"] < info2["train"][0]["objective"]
    assert T.find_last_checkpoint(str(tmp_path / T.exp_dir_name("senet50-ferplus", "emovoxceleb-student", "hot-cross-ent", 1, 8, "max", 2))) == 2
    # 'continue': a third epoch resumes from net-epoch-2 (parameters AND momentum) ...
    p3, info3 = T.run_distillation(imdb, get_batch, root=str(tmp_path), numEpochs=3, **common)
    assert len(info3["train"]) == 3 and info3["train"][:2] == info2["train"]
    # ... and equals three uninterrupted epochs (up to the summation order of the split-K filter-gradient atomics)
    p3b, info3b = T.run_distillation(imdb, get_batch, root=str(tmp_path / "fresh"), numEpochs=3, **common)
    assert abs(info3["train"][2]["objective"] - info3b["train"][2]["objective"]) < 1e-3 * info3b["train"][2]["objective"]
    for k in p3:
        if deterministic:
            assert np.array_equal(p3[k], p3b[k]), k
            continue
        if not (k.endswith("f") or k.endswith("m") or k.endswith("x")):
            continue  # zero-initialised biases hold only lr * (chaotic) gradient after three epochs
        # (BN moving-average moments of an 8-sample batch amplify the atomics' summation-order noise the most)
        assert rel_err(p3[k], p3b[k]) < (1e-2 if k.endswith("x") else 1e-3), k
    if deterministic:
        assert info3["train"][2]["objective"] == info3b["train"][2]["objective"]
