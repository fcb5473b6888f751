"""Data-parallel check of the graph-level C ABI on >= 2 GPUs (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/tools/dp_check.py
Every rank runs xemo_distill_step with a communicator created by the library (ncclAllReduce inside the captured step) on
its own inputs.  Checked: (1) the all-reduced gradient equals the sum of the ranks' local gradients, (2) all ranks end the
steps with BIT-IDENTICAL parameters and momentum, (3) the objective each rank reports is its own local one.  BN moments: the batch moments are summed over the ranks too
(one more small all-reduce in the captured step), so the moving averages agree bit for bit as well."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcncrossmodalemotions_b200 import zoo  # noqa: E402
from mcncrossmodalemotions_b200.net import Comm, DistillStep, StudentNet  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, width = 4, 100
    tp, sp = zoo.teacher_init("senet50"), zoo.student_init()
    rng = np.random.default_rng(50 + rank)
    faces = rng.integers(0, 256, (48, 48, n), dtype=np.uint8)
    spec = rng.standard_normal((512, width, 1, n)).astype(np.float32)
    step = DistillStep(tp, sp, n, width, device=local)
    step.comm = Comm.from_torch(step.ctx)
    step.student.set_hyper(lr=1e-3, batch_size=n * world)
    step.teacher.set_input(faces)
    step.student.set_input(spec)
    step.step()
    g_sum = step.student.export_grads()
    m = step.student.metrics()
    # the same first step without a communicator: the local gradient
    solo = DistillStep(tp, sp, n, width, device=local)
    solo.student.set_hyper(lr=1e-3, batch_size=n * world)
    solo.teacher.set_input(faces)
    solo.student.set_input(spec)
    solo.step()
    g_loc = solo.student.export_grads()
    assert abs(solo.student.metrics()["objective"] - m["objective"]) <= 1e-5 * abs(m["objective"])
    worst = 0.0
    for k in sorted(g_loc):
        if k.endswith("x") or (k.endswith("b") and not k.startswith("bn") and k != "fc8b"):
            continue      # batch moments; conv biases ahead of train-mode BN (zero gradient: cancellation noise)
        t = torch.from_numpy(np.ascontiguousarray(g_loc[k])).cuda()
        dist.all_reduce(t)
        ref = t.cpu().numpy()
        d = np.abs(ref).max()
        if d < 1e-7:
            continue
        err = float(np.abs(g_sum[k] - ref).max() / d)
        worst = max(worst, err)
        assert err < 1e-4, (k, err)       # (filter-gradient atomics: two runs of the same step differ in the last bits)
    for _ in range(3):
        step.step()
    step.sync()
    for what, tensors in (("params", step.student.export_params()), ("momentum", step.student.export_momentum())):
        for k in sorted(tensors):
            # (BN moments included: the batch moments are summed across the ranks like any other derivative, as MatConvNet's
            # parameter server does, so every rank holds the same moving averages)
            t = torch.from_numpy(np.ascontiguousarray(tensors[k])).cuda()
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), "%s %s differs between ranks" % (what, k)
    dist.barrier()
    kernels = step.num_kernels()
    for obj in (step.student, step.teacher, solo.student, solo.teacher, step.comm):     # networks before the communicator
        obj.close()
    if rank == 0:
        print("dp_check ok: world %d, all-reduced gradient == sum of local gradients (worst rel %.1e), parameters bit-identical "
              "across ranks after 4 steps, %d kernels per step" % (world, worst, kernels))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
