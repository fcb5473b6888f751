"""Achieved HBM GB/s of the memory-bound kernels of one distillation step: algorithmic bytes (tensor sizes read + written,
fp16 activations, per-GPU batch 256, 512x300 spectrograms) over the per-launch duration of an ncu launch list
(`--metrics gpu__time_duration.sum --clock-control none`: serialised, cold-cache).
    python tools/hbm_table.py profiles/r02_launches_step.csv [peak_GBs] > profiles/r02_hbm_kernels.txt
Round-2 kernel sequence: SE by linearity on the 56x56 / 28x28 stages (their squeeze reads the 3x3 output, C/4 channels; no
excite pass), squeeze -> gate -> excite on the 14x14 / 7x7 stages; pooled tensors of the student's first layer carry a
128-channel pitch.  (The round-1 sequence: git history of this file.)"""
import csv
import json
import os
import sys

path = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6554.2) \
    if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6554.2
N = 256
rows = list(csv.reader(open(path)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[h]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
seq = []
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] == "ns" else v if r[ui] == "us" else v * 1e3
    seq.append((r[ki].split("(")[0].replace("void ", "").replace("xemo::", ""), us))
starts = [i for i, (n, _) in enumerate(seq) if "face_u8" in n]
st = starts[1] if len(starts) > 1 else starts[0]
step = seq[st:st + 188]
MB = 1e6
act = lambda h_, w_, c: N * h_ * w_ * c * 2 / MB          # fp16 NHWC tensor of the batch, MB
stage_hw = {2: 56, 3: 28, 4: 14, 5: 7}
stage_c = {2: 256, 3: 512, 4: 1024, 5: 2048}
stage_blocks = [2] * 3 + [3] * 4 + [4] * 6 + [5] * 3
excite_blocks = [4] * 6 + [5] * 3
student_bn = [(62, 36, 256), (30, 17, 384), (30, 17, 256), (30, 17, 256), (1, 8, 4096), (1, 1, 1024)]   # bn2..bn7 inputs
out = []
sq = ex = bs = br = ba = 0
for name, us in step:
    b, what = None, ""
    if name.startswith("face_u8"):
        b, what = N * 48 * 48 / MB + act(224, 112, 32), "uint8 faces -> row-im2col [N][224][112][32]"
    elif name.startswith("maxpool_fwd_h2_kernel<0, 3, 3, 0>"):
        b, what = act(112, 112, 64) + act(56, 56, 64), "teacher pool1"
    elif name.startswith("se_squeeze"):
        s = stage_blocks[sq]; sq += 1
        lin = s <= 3                                      # by linearity: the squeeze reads t2 (C / 4 channels)
        b, what = act(stage_hw[s], stage_hw[s], stage_c[s] // (4 if lin else 1)), "SE squeeze, stage %d%s" % (s, " (of the 3x3 output)" if lin else "")
    elif name.startswith("se_excite"):
        s = excite_blocks[ex]; ex += 1
        b, what = 3 * act(stage_hw[s], stage_hw[s], stage_c[s]), "SE excite + shortcut + ReLU, stage %d" % s
    elif name.startswith("spec_s2d"):
        b, what = N * 512 * 300 * 4 / MB + act(257, 148, 16), "spectrogram -> space-to-depth"
    elif name.startswith("stem_autocorr"):
        b, what = act(257, 148, 16), "patch autocorrelation (mma.sync; compute-bound)"
    elif name.startswith("maxpool_fwd_h2_kernel<1, 3, 3, 1>") and us > 300:
        b, what = act(254, 148, 96) + 2.5 * act(126, 73, 128), "student pool1: BN+ReLU folded, y + arg-max + winner (128-channel pitch)"
    elif name.startswith("maxpool_fwd_h2_kernel<1, 3, 3, 1>"):
        b, what = act(62, 36, 256) + 1.5 * act(30, 17, 256), "student pool2"
    elif name.startswith("bn_stats"):
        g = student_bn[bs]; bs += 1
        b, what = act(*g), "BN statistics, %dx%dx%d" % g
    elif name.startswith("bn_bwd_reduce"):
        g = student_bn[::-1][br]; br += 1
        b, what = 2 * act(*g), "BN backward reduce, %dx%dx%d" % g
    elif name.startswith("bn_bwd_apply"):
        g = student_bn[::-1][ba]; ba += 1
        b, what = 3 * act(*g), "BN backward apply, %dx%dx%d" % g
    elif name.startswith("stem_pool_bn_reduce"):
        b, what = 3 * act(126, 73, 128), "stem BN reductions + mask at the pooled resolution"
    elif name.startswith("maxpool_bwd_3x3s2") and us > 300:
        b, what = 1.5 * act(126, 73, 128) + act(254, 148, 96), "student pool1 backward"
    elif name.startswith("maxpool_bwd_3x3s2"):
        b, what = 1.5 * act(30, 17, 256) + act(62, 36, 256), "student pool2 backward"
    elif name.startswith("sgd_momentum"):
        b, what = 16.63e6 * 26 / MB, "SGD-momentum over the flat parameter buffer (+ fp16 mirror)"
    if b is not None and us > 15:
        out.append((name[:34], what, b, us, b / us * 1e3))   # MB / us = TB/s; x 1000 = GB/s
print("kernel                             | role                                                      |   MB   |   us   |  GB/s | of %.0f" % peak)
for name, what, b, us, gbs in out:
    print("%-34s | %-57s | %6.0f | %6.1f | %5.0f | %.2f" % (name, what, b, us, gbs, gbs / peak))
