"""Checkpoints for the two networks of the OCR stage.

EasyOCR loads `craft_mlt_25k.pth` (detector) and `english_g2.pth` (recogniser) from ~/.EasyOCR/model
(easyocr/easyocr.py Reader.__init__, reached from pipeline_demo/extractor/enhanced_extractor.py:153).  Neither file
exists in this image and there is no network, so this module provides

  * `find_checkpoints()` / `load_pth()`  : use the genuine files when they are present (keys are the upstream ones,
    with the `module.` prefix of DataParallel checkpoints stripped, as easyocr.detection.copyStateDict does);
  * `random_craft_state(seed)` / `random_crnn_state(seed)` : seeded random state dicts with identical key names and
    shapes (He-initialised convolutions, non-trivial BatchNorm statistics) so every kernel can be parity-tested and
    benchmarked without the real files.  The generator is NumPy's PCG64 -> bit-identical on every machine;
  * `calibrated_craft_state(seed)` : the random detector with its decoder tail (upconv4 + conv_cls, 54 k parameters)
    replaced by one fitted on synthetic pages (bbocr_b200/data/craft_probe.npz, produced by tools/fit_craft_probe.py)
    so that the network's score maps light up on synthetic text and `getDetBoxes` sees a realistic number of
    components.  Everything in front of that tail stays the seeded random state.

State dicts are plain {name: float32 ndarray}.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512]   # conv5_3 unused by CRAFT


def _conv(rng, sd, name, cout, cin, kh, kw, bias=True, gain=2.0):
    fan_in = cin * kh * kw
    sd[name + ".weight"] = (rng.standard_normal((cout, cin, kh, kw)) * np.sqrt(gain / fan_in)).astype(np.float32)
    if bias:
        sd[name + ".bias"] = (rng.standard_normal(cout) * 0.05).astype(np.float32)


def _bn(rng, sd, name, c):
    sd[name + ".weight"] = rng.uniform(0.8, 1.2, c).astype(np.float32)
    sd[name + ".bias"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
    sd[name + ".running_mean"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
    sd[name + ".running_var"] = rng.uniform(0.7, 1.3, c).astype(np.float32)


def random_craft_state(seed: int = 1234) -> "OrderedDict[str, np.ndarray]":
    """Keys of easyocr/craft.py::CRAFT().state_dict() (minus num_batches_tracked)."""
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    idx, cin = 0, 3
    slice_of = lambda i: 1 if i < 12 else 2 if i < 19 else 3 if i < 29 else 4
    for v in VGG_CFG:
        if v == "M":
            idx += 1
            continue
        s = slice_of(idx)
        _conv(rng, sd, f"basenet.slice{s}.{idx}", v, cin, 3, 3)
        _bn(rng, sd, f"basenet.slice{slice_of(idx + 1)}.{idx + 1}", v)
        idx += 3
        cin = v
    _conv(rng, sd, "basenet.slice5.1", 1024, 512, 3, 3)
    _conv(rng, sd, "basenet.slice5.2", 1024, 1024, 1, 1, gain=1.0)
    for n, (i, m, o) in enumerate([(1024, 512, 256), (512, 256, 128), (256, 128, 64), (128, 64, 32)], start=1):
        _conv(rng, sd, f"upconv{n}.conv.0", m, i + m, 1, 1)
        _bn(rng, sd, f"upconv{n}.conv.1", m)
        _conv(rng, sd, f"upconv{n}.conv.3", o, m, 3, 3)
        _bn(rng, sd, f"upconv{n}.conv.4", o)
    _conv(rng, sd, "conv_cls.0", 32, 32, 3, 3)
    _conv(rng, sd, "conv_cls.2", 32, 32, 3, 3)
    _conv(rng, sd, "conv_cls.4", 16, 32, 3, 3)
    _conv(rng, sd, "conv_cls.6", 16, 16, 1, 1)
    _conv(rng, sd, "conv_cls.8", 2, 16, 1, 1, gain=1.0)
    return sd


def calibrated_craft_state(seed: int = 1234):
    sd = random_craft_state(seed)
    p = os.path.join(_DATA, "craft_probe.npz")
    if os.path.exists(p):
        z = np.load(p)
        if int(z["seed"]) == seed:
            for k in z.files:
                if k != "seed":
                    assert sd[k].shape == z[k].shape, k
                    sd[k] = z[k].astype(np.float32)
    return sd


def random_crnn_state(seed: int = 4321, num_class: int = 97, hidden: int = 256) -> "OrderedDict[str, np.ndarray]":
    """Keys of easyocr/model/vgg_model.py::Model(1, 256, 256, 97).state_dict()."""
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    p = "FeatureExtraction.ConvNet."
    _conv(rng, sd, p + "0", 32, 1, 3, 3)
    _conv(rng, sd, p + "3", 64, 32, 3, 3)
    _conv(rng, sd, p + "6", 128, 64, 3, 3)
    _conv(rng, sd, p + "8", 128, 128, 3, 3)
    _conv(rng, sd, p + "11", 256, 128, 3, 3, bias=False)
    _bn(rng, sd, p + "12", 256)
    _conv(rng, sd, p + "14", 256, 256, 3, 3, bias=False)
    _bn(rng, sd, p + "15", 256)
    _conv(rng, sd, p + "18", 256, 256, 2, 2)
    k = 1.0 / np.sqrt(hidden)
    for layer in (0, 1):
        q = f"SequenceModeling.{layer}."
        for sfx in ("", "_reverse"):
            sd[q + "rnn.weight_ih_l0" + sfx] = rng.uniform(-k, k, (4 * hidden, 256)).astype(np.float32)
            sd[q + "rnn.weight_hh_l0" + sfx] = rng.uniform(-k, k, (4 * hidden, hidden)).astype(np.float32)
            sd[q + "rnn.bias_ih_l0" + sfx] = rng.uniform(-k, k, 4 * hidden).astype(np.float32)
            sd[q + "rnn.bias_hh_l0" + sfx] = rng.uniform(-k, k, 4 * hidden).astype(np.float32)
        kl = 1.0 / np.sqrt(2 * hidden)
        sd[q + "linear.weight"] = rng.uniform(-kl, kl, (hidden, 2 * hidden)).astype(np.float32)
        sd[q + "linear.bias"] = rng.uniform(-kl, kl, hidden).astype(np.float32)
    sd["Prediction.weight"] = (rng.uniform(-k, k, (num_class, hidden)) * 6.0).astype(np.float32)
    sd["Prediction.bias"] = rng.uniform(-k, k, num_class).astype(np.float32)
    return sd


def calibrated_crnn_state(seed: int = 4321):
    """Random recogniser with its Prediction read-out fitted on synthetic lines (tools/fit_crnn_probe.py)."""
    sd = random_crnn_state(seed)
    p = os.path.join(_DATA, "crnn_probe.npz")
    if os.path.exists(p):
        z = np.load(p)
        if int(z["seed"]) == seed:
            for k in z.files:
                if k != "seed":
                    assert sd[k].shape == z[k].shape, k
                    sd[k] = z[k].astype(np.float32)
    return sd


# ----------------------------------------------------------------------------------------------------------------------

def find_checkpoints(model_storage_directory: str | None = None):
    """-> (craft_path | None, crnn_path | None): EasyOCR's MODULE_PATH/model layout."""
    roots = [model_storage_directory] if model_storage_directory else []
    roots += [os.environ.get("EASYOCR_MODULE_PATH") and os.path.join(os.environ["EASYOCR_MODULE_PATH"], "model"),
              os.path.expanduser("~/.EasyOCR/model")]
    craft = crnn = None
    for r in roots:
        if not r:
            continue
        a, b = os.path.join(r, "craft_mlt_25k.pth"), os.path.join(r, "english_g2.pth")
        craft = craft or (a if os.path.exists(a) else None)
        crnn = crnn or (b if os.path.exists(b) else None)
    return craft, crnn


def load_pth(path: str):
    import torch
    raw = torch.load(path, map_location="cpu", weights_only=True)
    sd = OrderedDict()
    for k, v in raw.items():
        if k.startswith("module."):
            k = k[len("module."):]
        if k.endswith("num_batches_tracked"):
            continue
        sd[k] = v.detach().cpu().numpy().astype(np.float32)
    return sd


def to_torch_state(sd):
    """Plain ndarray dict -> tensors, adding the BatchNorm counters torch's strict loading expects."""
    import torch
    out = OrderedDict()
    for k, v in sd.items():
        out[k] = torch.from_numpy(np.ascontiguousarray(v))
        if k.endswith("running_var"):
            out[k[:-len("running_var")] + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return out
