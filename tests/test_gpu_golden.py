"""The CUDA path (through the C ABI) against the committed golden vectors of tests/golden/ -- no oracle code runs
here: inputs are the fixture's own arrays or the seeded synthetic generators, targets are the stored fp64 results."""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_boundary_operators_against_golden():
    from mcncrossmodalemotions_b200 import vl_nn as V

    g = np.load(os.path.join(G, "ops.npz"))
    f32 = lambda k: g[k].astype(np.float32)
    assert rel_err(V.vl_nnconv(f32("conv_x"), f32("conv_f"), f32("conv_b"), pad=(1, 1, 0, 1), stride=(2, 2)), g["conv_y"]) < 1e-3
    dx, df, db = V.vl_nnconv(f32("conv_x"), f32("conv_f"), f32("conv_b"), f32("conv_dy"), pad=(1, 1, 0, 1), stride=(2, 2))
    assert rel_err(dx, g["conv_dx"]) < 1e-3 and rel_err(df, g["conv_df"]) < 1e-3 and rel_err(db, g["conv_db"]) < 2e-5
    y, idx = V.vl_nnpool(f32("pool_x"), (3, 3), pad=(0, 1, 0, 1), stride=2, method="max", return_index=True)
    assert np.array_equal(y, g["pool_y"]) and np.array_equal(idx, g["pool_idx"])          # bit-exact pooling indices
    assert rel_err(V.vl_nnpool(f32("pool_x"), (3, 3), f32("pool_dy"), pad=(0, 1, 0, 1), stride=2, method="max"), g["pool_dx"]) < 2e-5
    assert rel_err(V.vl_nnpool(f32("pool_x"), (2, 3), pad=(1, 0, 1, 1), stride=(2, 1), method="avg"), g["avg_y"]) < 2e-5
    y, mom = V.vl_nnbnorm(f32("bn_x"), f32("bn_g"), f32("bn_b"), epsilon=1e-5)
    assert rel_err(y, g["bn_y"]) < 2e-5 and rel_err(mom, g["bn_mom"]) < 2e-5
    dx, dg, db, _ = V.vl_nnbnorm(f32("bn_x"), f32("bn_g"), f32("bn_b"), f32("bn_dy"), epsilon=1e-5)
    assert rel_err(dx, g["bn_dx"]) < 1e-4 and rel_err(dg, g["bn_dg"]) < 1e-4 and rel_err(db, g["bn_db"]) < 1e-4
    assert abs(V.vl_nnsoftmaxceloss(f32("loss_x"), f32("loss_t"), temperature=2.0, logitTargets=True) - float(g["loss_y"])) < 1e-5 * float(g["loss_y"])
    assert rel_err(V.vl_nnsoftmaxceloss(f32("loss_x"), f32("loss_t"), 1.0, temperature=2.0, logitTargets=True), g["loss_dx"]) < 2e-5


def test_networks_against_golden():
    from mcncrossmodalemotions_b200 import zoo
    from mcncrossmodalemotions_b200.programs import StudentProgram, TeacherProgram

    g = np.load(os.path.join(G, "nets.npz"))
    rng_faces = np.random.default_rng(0).uniform(0, 255, (224, 224, 3, 2)).astype(np.float32) - np.array(zoo.AVERAGE_IMAGE, np.float32).reshape(1, 1, 3, 1)
    for arch in ("resnet50", "senet50"):
        got = TeacherProgram(zoo.teacher_init(arch), 2).forward(rng_faces)
        assert rel_err(got, g["teacher_%s_logits" % arch]) < 1e-3
    u8 = np.random.default_rng(0).integers(0, 256, (48, 48, 2), dtype=np.uint8)
    got = TeacherProgram(zoo.teacher_init("senet50"), 2, input_mode="u8").forward(u8)
    assert rel_err(got, g["teacher_senet50_logits_48"]) < 1e-3
    # one training step of the freshly initialised student: continuous quantities at 1e-3
    n = 4
    s = np.random.default_rng(1).standard_normal((512, 100, 1, n))
    spec = ((s - s.mean(axis=1, keepdims=True)) / s.std(axis=1, ddof=1, keepdims=True)).astype(np.float32)
    tgt = (3.0 * np.random.default_rng(2).standard_normal((1, 1, 8, n))).astype(np.float32)
    prog = StudentProgram(zoo.student_init(), n, 100)
    prog.set_hyper(lr=1e-4, batch_size=n)
    prog.reset_metrics()
    prog.set_input(spec, tgt)
    prog.grad_step()
    m, grads = prog.metrics(), prog.export_grads()
    assert abs(m["objective"] - float(g["student_step_objective"])) < 1e-3 * float(g["student_step_objective"])
    assert m["classerror"] == float(g["student_step_classerror"])
    for k, tol in (("bn1x", 1e-3), ("bn4x", 1e-3), ("bn7x", 5e-2)):   # bn7 normalises over only N = 4 rows here
        assert rel_err(grads[k], g["student_step_" + k]) < tol, k
    for k in ("fc8f", "fc6f", "conv3f", "conv1f"):
        assert abs(np.linalg.norm(grads[k]) - float(g["student_step_gradnorm_" + k])) < 0.1 * float(g["student_step_gradnorm_" + k]), k
