#!/usr/bin/env python
"""Fit the Prediction layer (256 -> 97 linear read-out) of the seeded random CRNN on synthetic text lines.

There are no trained recogniser weights in this image (no network).  With a purely random read-out the CRNN emits
the same two or three classes at every time step, so greedy CTC strings and confidences would not depend on the crop.
A ridge-regressed read-out of the random conv + BiLSTM features onto frame-wise character targets (known glyph
positions of rendered lines) gives crop-dependent strings and a spread of confidences, so that the contrast-retry
branch (conf < 0.1) and the CTC collapse are exercised realistically.  Everything in front of the read-out stays the
seeded random state.  Output: bbocr_b200/data/crnn_probe.npz (float16).  CPU only, about a minute.
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from PIL import Image, ImageDraw

from bbocr_b200 import synth, weights
from oracle import easyocr_restated as E

SEED = 4321


def render_line(rng):
    n = int(rng.integers(2, 22))
    words = []
    while sum(len(w) + 1 for w in words) < n:
        words.append(synth._WORDS[int(rng.integers(len(synth._WORDS)))])
    text = " ".join(words)[:n].strip() or "a"
    if rng.random() < 0.3:
        text = text.upper()
    f = synth._font(int(rng.integers(34, 50)))
    probe = ImageDraw.Draw(Image.new("L", (8, 8)))
    xs = [8 + probe.textlength(text[:i], font=f) for i in range(len(text) + 1)]
    w = int(xs[-1]) + 12
    img = Image.new("L", (w, 64), int(rng.integers(190, 246)))
    ImageDraw.Draw(img).text((8, 6), text, fill=int(rng.integers(0, 70)), font=f)
    a = np.asarray(img).astype(np.float32) + rng.normal(0, 3, (64, w)).astype(np.float32)
    return np.clip(a, 0, 255).astype(np.uint8), text, xs


def main():
    rng = np.random.default_rng(7)
    crnn = E.CRNN()
    crnn.load_state_dict(weights.to_torch_state(weights.random_crnn_state(SEED)))
    crnn.eval()
    feats, targets = [], []
    for i in range(500):
        crop, text, xs = render_line(rng)
        Wm = int(np.ceil(crop.shape[1] / 64)) * 64
        x = E.align_collate_one(crop, Wm)[None, None]
        with torch.no_grad():
            v = crnn.FeatureExtraction(torch.from_numpy(x))
            v = crnn.AdaptiveAvgPool(v.permute(0, 3, 1, 2)).squeeze(3)
            c = crnn.SequenceModeling(v)[0].numpy()           # T x 256
        T = c.shape[0]
        tgt = np.zeros(T, np.int64)
        for t in range(T):
            xc = 4 * t + 4.0                                  # centre of the step's receptive field (input px)
            for k, ch in enumerate(text):
                a, b = xs[k], xs[k + 1]
                m = 0.2 * (b - a)
                if a + m <= xc <= b - m and ch != " ":
                    tgt[t] = E.CHARACTERS.index(ch) + 1
                    break
        feats.append(c)
        targets.append(tgt)
    X = np.concatenate(feats).astype(np.float64)
    y = np.concatenate(targets)
    Y = np.full((len(y), 97), -1.0)
    Y[np.arange(len(y)), y] = 1.0
    A = np.concatenate([X, np.ones((len(X), 1))], 1)
    G = A.T @ A + 1e-2 * len(X) * np.eye(257) * 1e-3
    Wt = np.linalg.solve(G, A.T @ Y)                          # 257 x 97
    pred = (A @ Wt).argmax(1)
    print("frame accuracy", (pred == y).mean(), "non-blank recall", (pred[y > 0] == y[y > 0]).mean())
    scale = 36.0                                              # sharpen the soft-max so confidences straddle 0.1
    np.savez_compressed(os.path.join(os.path.dirname(weights.__file__), "data", "crnn_probe.npz"), seed=np.int64(SEED),
                        **{"Prediction.weight": (Wt[:256].T * scale).astype(np.float16),
                           "Prediction.bias": (Wt[256] * scale).astype(np.float16)})
    print("saved")


if __name__ == "__main__":
    main()
