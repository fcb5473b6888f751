"""Training driver: the parts of cnn_train_dag the reference relies on (emoVoxCeleb/run_distillation.m:170-182) --
epochs over a (sub-sampled) training set, per-epoch learning rate, SGD-momentum with weight decay, validation pass,
running objective / classerror / per-class accuracy (extractStats, run_distillation.m:186-207), checkpoint per epoch
and 'continue' (resume from the latest) -- on top of the graph-level C ABI (net.StudentNet: xemo_student_train_step /
xemo_sgd_step; the gradient sum across ranks is ncclAllReduce inside the library), plus `run_distillation`, the option
surface of the reference driver (run_distillation.m:71-90)."""
from __future__ import annotations

import glob
import json
import os
import socket
import time

import numpy as np

from . import zoo
from .net import Comm, StudentNet

EMOTIONS = ["neutral", "happiness", "surprise", "sadness", "anger", "disgust", "fear", "contempt"]  # FER+ order


def learning_rate_schedule(num_epochs=300, hi=-4.0, lo=-5.0):
    """opts.learningRate = logspace(-4, -5, numEpochs) (run_distillation.m:87)."""
    return np.logspace(hi, lo, num_epochs)


def exp_dir_name(teacher, student, loss_type, num_seconds, num_pred, aggregator, temperature, from_scratch=True):
    """run_distillation.m:95-104."""
    s = "%s-%s" % (student, loss_type) + ("-scratch" if from_scratch else "")
    name = "voxceleb-%s-%s-%dsec-%demo-agg-%s" % (teacher, s, num_seconds, num_pred, aggregator)
    return name + ("-temp%d" % temperature if loss_type == "hot-cross-ent" else "")


def extract_stats(metrics, num_samples):
    """extractStats (run_distillation.m:186-207): objective / classerror averages, meanAcc, per-emotion accuracy and
    population from the ErrorStats counters."""
    count = np.maximum(metrics["count"], 0)
    acc = np.where(count > 0, metrics["correct"] / np.maximum(count, 1), 0.0)
    stats = {"objective": metrics["objective"] / max(num_samples, 1), "classerror": metrics["classerror"] / max(num_samples, 1),
             "meanAcc": float(acc.mean())}
    pop = count / max(count.sum(), 1)
    for j, name in enumerate(EMOTIONS[: len(acc)]):
        stats[name] = float(acc[j])
        stats[name + "Pop"] = float(pop[j])
    return stats


def find_last_checkpoint(exp_dir):
    eps = [int(os.path.basename(f)[len("net-epoch-"):-4]) for f in glob.glob(os.path.join(exp_dir, "net-epoch-*.npz"))]
    return max(eps) if eps else 0


def save_checkpoint(exp_dir, epoch, program, stats):
    params = program.export_params()
    momentum = program.export_momentum()
    np.savez(os.path.join(exp_dir, "net-epoch-%d.npz" % epoch), **{"p:" + k: v for k, v in params.items()},
             **{"m:" + k: v for k, v in momentum.items()}, stats=json.dumps(stats))


class LossScaler:
    """Dynamic loss scale for the fp16 activation-gradient chain (the reference trains in `single` and has no such knob).
    A non-finite gradient -- detected on the device, where the guarded update has already skipped that step -- halves the
    scale; `growth_interval` consecutive clean steps double it again, within [min_scale, the initial scale].  Under data
    parallelism every rank sees the same flag (the guard scans the all-reduced gradient), so the ranks move in step."""

    def __init__(self, initial=1024.0, growth_interval=2000, min_scale=1.0):
        self.initial = self.scale = float(initial)
        self.growth_interval, self.min_scale = int(growth_interval), float(min_scale)
        self.clean = 0
        self.overflows = 0

    def update(self, overflow):
        """Feed one step's flag; returns the scale to use from the next step on."""
        if overflow:
            self.overflows += 1
            self.clean = 0
            self.scale = max(self.scale / 2.0, self.min_scale)
        else:
            self.clean += 1
            if self.clean >= self.growth_interval and self.scale < self.initial:
                self.scale = min(self.scale * 2.0, self.initial)
                self.clean = 0
        return self.scale


def apply_loss_scale(prog, scaler, overflow, log=print):
    """One step of the policy on a StudentNet-like object (`grad_scale`, `set_grad_scale`)."""
    new = scaler.update(overflow)
    if new != prog.grad_scale:
        log("loss scale %g -> %g (%s)" % (prog.grad_scale, new, "non-finite gradient, step skipped" if overflow else "clean run"))
        prog.set_grad_scale(new)
    return new


def load_checkpoint(exp_dir, epoch):
    z = np.load(os.path.join(exp_dir, "net-epoch-%d.npz" % epoch), allow_pickle=False)
    params = {k[2:]: z[k] for k in z.files if k.startswith("p:")}
    momentum = {k[2:]: z[k] for k in z.files if k.startswith("m:")}
    return params, momentum, json.loads(str(z["stats"]))


def cnn_train_dag(params, imdb, get_batch, *, learning_rate, batch_size=64, num_epochs=300, train=None, val=None, cont=True,
                  exp_dir=None, width=400, epoch_size=None, momentum=0.9, weight_decay=5e-4, device=0, world=1, rank=0, seed=0,
                  max_steps_per_epoch=None, log=print, loss_type="hot-cross-ent"):
    """Train the student.  `params`: zoo parameter dict; `imdb`: anything `get_batch(imdb, indices)` understands;
    get_batch returns the dict of batch.get_batch ('data', 'logitTarget').  Returns (params, info)."""
    train = np.asarray(train if train is not None else [], np.int64)
    val = np.asarray(val if val is not None else [], np.int64)
    start = 0
    info = {"train": [], "val": []}
    mom = None
    if exp_dir:
        os.makedirs(exp_dir, exist_ok=True)
        if cont:
            start = find_last_checkpoint(exp_dir)
            if start:
                params, mom, info = load_checkpoint(exp_dir, start)
                log("resuming from epoch %d" % start)
    per_rank = batch_size // world
    prog = StudentNet(params, per_rank, width, device=device, loss_type=loss_type)
    tkey = "maxLabel" if loss_type == "softmaxlog" else "logitTarget"   # the loss layer's second input (emoVoxZoo.m:137-157)
    if mom is not None:
        prog.load_momentum(mom)
    comm = Comm.from_torch(prog.ctx) if world > 1 else None   # one process per GPU, torch.distributed only hands out the NCCL id
    scaler = LossScaler(prog.grad_scale)
    for epoch in range(start, num_epochs):
        rng = np.random.default_rng([seed, epoch])   # per-epoch stream: a resumed run draws the same permutations
        lr = float(learning_rate[min(epoch, len(learning_rate) - 1)])
        prog.set_hyper(lr=lr, momentum=momentum, weight_decay=weight_decay, batch_size=batch_size)
        order = rng.permutation(train)
        if epoch_size:
            order = order[: int(epoch_size)]   # 'epochSize': a random subset of the training set per epoch
        prog.reset_metrics()
        t0, seen, obj, err = time.time(), 0, 0.0, 0.0
        # (a trailing partial batch is dropped: train-mode BN statistics and the captured graphs are tied to the batch size;
        # upstream cnn_train_dag would run it as a smaller batch)
        steps = len(order) // batch_size
        if max_steps_per_epoch:
            steps = min(steps, max_steps_per_epoch)
        for it in range(steps):
            idx = order[it * batch_size : (it + 1) * batch_size][rank::world]   # labindex:numlabs:end
            inputs = get_batch(imdb, idx)
            prog.train_step(inputs["data"], inputs[tkey], comm, weights=inputs.get("instanceWeights"))
            m = prog.metrics()   # objective / classerror of this batch; class counters accumulate
            apply_loss_scale(prog, scaler, m["nonfinite_grad"], log)
            if m["nonfinite_grad"]:
                continue             # (the update was skipped on the device; its objective is not a number either)
            obj += m["objective"]; err += m["classerror"]; seen += len(idx)
        m = prog.metrics()
        tr = extract_stats(dict(objective=obj, classerror=err, correct=m["correct"], count=m["count"]), seen)
        tr["speed_hz"] = seen / max(time.time() - t0, 1e-9)
        info["train"].append(tr)
        vs = None
        if len(val):
            vs = evaluate(prog, imdb, get_batch, val, per_rank, tkey)
            info["val"].append(vs)
        log("epoch %d lr %.3g train obj %.4f err %.3f%s" % (epoch + 1, lr, tr["objective"], tr["classerror"],
                                                          "" if vs is None else " | val err %.3f" % vs["classerror"]))
        if exp_dir and rank == 0:
            save_checkpoint(exp_dir, epoch + 1, prog, info)
    return prog.export_params(), info


def evaluate(prog, imdb, get_batch, indices, batch, tkey="logitTarget"):
    """Validation pass: forward in test mode, class error against the arg-max teacher label.  The last, partial batch is
    processed too (padded with repeats of its last sample; the padding rows are discarded) -- test-mode BN makes every
    sample independent of its batch."""
    wrong = total = 0
    for it in range((len(indices) + batch - 1) // batch):
        idx = np.asarray(indices[it * batch : (it + 1) * batch])
        n = len(idx)
        if n < batch:
            idx = np.concatenate([idx, np.repeat(idx[-1:], batch - n)])
        inputs = get_batch(imdb, idx)
        pred = prog.forward(inputs["data"], "test")[:n]
        if tkey == "maxLabel":
            label = np.asarray(inputs["maxLabel"]).reshape(-1).astype(np.int64)[:n] - 1
        else:
            label = inputs["logitTarget"].reshape(pred.shape[1], -1).argmax(axis=0)[:n]
        wrong += int((pred.argmax(axis=1) != label).sum())
        total += n
    return {"classerror": wrong / max(total, 1), "num": total}


def store_meta_info(opts, exp_dir):
    """storeMetaInfo (run_distillation.m:227-240): options + hostname next to the checkpoints."""
    stamp = time.strftime("%d-%b-%Y_%H-%M-%S")
    txt = "server: %s\n" % socket.gethostname() + "".join("%s: %s\n" % (k, v) for k, v in sorted(opts.items()))
    with open(os.path.join(exp_dir, "meta-%s.txt" % stamp), "w") as f:
        f.write(txt)
    with open(os.path.join(exp_dir, "meta-%s.json" % stamp), "w") as f:
        json.dump({k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in opts.items()}, f, default=str)


def run_distillation(imdb, get_batch, root="data/xEmo18", **overrides):
    """run_distillation (emoVoxCeleb/run_distillation.m): same option names and defaults; `imdb` / `get_batch` stand
    for fetch_emovoxceleb_imdb + getBatchFn (dataset access is out of scope of the hot path)."""
    opts = dict(gpus=[0], cont=True, miniVal=0.2, numSeconds=4, batchSize=64, numEpochs=300, numPredEmotions=8, fromScratch=True,
                logitAggregator="max", datasetName="voxceleb", teacher="senet50-ferplus", student="emovoxceleb-student",
                lossType="hot-cross-ent", temperature=2, fixedSegments=False, parameterServer="tmove", train=None, val=None)
    unknown = set(overrides) - set(opts) - {"miniEpochRatio", "learningRate", "max_steps_per_epoch"}
    if unknown:
        raise ValueError("unknown option(s): %s" % sorted(unknown))  # vl_argparse rejects unknown names
    extra = {k: overrides.pop(k) for k in ("miniEpochRatio", "learningRate", "max_steps_per_epoch") if k in overrides}
    opts.update(overrides)
    opts["miniEpochRatio"] = extra.get("miniEpochRatio", 0.05 * len(opts["gpus"]))
    opts["learningRate"] = extra.get("learningRate", learning_rate_schedule(opts["numEpochs"]))
    exp_dir = os.path.join(root, exp_dir_name(opts["teacher"], opts["student"], opts["lossType"], opts["numSeconds"],
                                              opts["numPredEmotions"], opts["logitAggregator"], opts["temperature"], opts["fromScratch"]))
    os.makedirs(exp_dir, exist_ok=True)
    net = zoo.emoVoxZoo(opts["student"], scratch=opts["fromScratch"], lossType=opts["lossType"], numSeconds=opts["numSeconds"],
                        numOutputs=opts["numPredEmotions"])
    train, val = np.asarray(opts["train"]), np.asarray(opts["val"])
    world, rank = 1, 0
    try:    # 'gpus', opts.gpus (run_distillation.m:179): one process per GPU under torch.distributed.run
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(), dist.get_rank()
    except ImportError:
        pass
    if opts["miniVal"] and len(val):
        val = np.random.default_rng(0).permutation(val)[: max(1, int(round(opts["miniVal"] * len(val))))]  # seeded val subsample
    store_meta_info(opts, exp_dir)
    return cnn_train_dag(net.params, imdb, get_batch, learning_rate=opts["learningRate"], batch_size=opts["batchSize"],
                         num_epochs=opts["numEpochs"], train=train, val=val, cont=opts["cont"], exp_dir=exp_dir,
                         width=100 * opts["numSeconds"], epoch_size=int(len(train) * opts["miniEpochRatio"]),
                         device=opts["gpus"][rank % len(opts["gpus"])], world=world, rank=rank, max_steps_per_epoch=extra.get("max_steps_per_epoch"), loss_type=opts["lossType"])
