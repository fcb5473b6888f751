"""Default-OFF options that were written without GPU time left in the round.  They are NOT selected
by `-m gpu` and skip themselves everywhere else unless asked for:
    XEMO_EXPERIMENTAL=1 python -m pytest tests -m gpu_experimental
on the GPU box, before switching any of them on."""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = [pytest.mark.gpu_experimental,
              pytest.mark.skipif(os.environ.get("XEMO_EXPERIMENTAL") != "1", reason="experimental GPU option: set XEMO_EXPERIMENTAL=1")]


@pytest.mark.parametrize("n", [8, 5])
def test_se_blocks_by_linearity_match_the_default_path_and_the_oracle(n):
    """XEMO_SE_LIN: squeeze(t2) -> gate (W3 folded in) -> expand convolution with the excite in its epilogue
    (conv_fprop_kernel<64, true>); n = 5 makes the 7 x 7 stage's 128-row tiles straddle up to four images."""
    from oracle import nets
    from mcncrossmodalemotions_b200.programs import TeacherProgram

    p = nets.teacher_init("senet50")
    x = nets.synth_faces(n)
    ref = nets.teacher_forward({k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in p.items()},
                               x.astype(np.float64), nets.TorchOps).reshape(8, n).T
    base = TeacherProgram(p, n, use_graph=False).forward(x)
    os.environ["XEMO_SE_LIN"] = "1"
    try:
        lin_prog = TeacherProgram(p, n, use_graph=False)
    finally:
        os.environ.pop("XEMO_SE_LIN")
    assert lin_prog.se_lin
    lin = lin_prog.forward(x)
    assert rel_err(base, ref) < 1e-3
    assert rel_err(lin, ref) < 1e-3
    assert rel_err(lin, base) < 1e-3
