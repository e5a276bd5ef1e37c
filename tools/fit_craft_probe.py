#!/usr/bin/env python
"""Fit the decoder tail (upconv4 + conv_cls) of the seeded random CRAFT so its score maps respond to synthetic text.

There are no trained detector weights in this image (no network).  The random VGG front-end still separates ink from
paper at relu2_2; training only the last decoder block and the classification head (54 k parameters, everything else
stays the seeded random state) on blob targets (synth.score_maps_for) gives score maps with a realistic number of
connected components for getDetBoxes and a realistic number of crops for the recogniser.
Output: bbocr_b200/data/craft_probe.npz (float16, ~110 kB).  CPU only, a few minutes.
"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cv2
import numpy as np
import torch
import torch.nn.functional as F

from bbocr_b200 import synth, weights
from oracle import easyocr_restated as E

SEED = 1234


def cached_inputs(craft, rgb):
    x, _ = E.canvas_tensor(rgb)
    with torch.no_grad():
        s = craft.basenet(x)
        y = craft.upconv1(torch.cat([s[0], s[1]], 1))
        for up, src in ((craft.upconv2, s[2]), (craft.upconv3, s[3])):
            y = F.interpolate(y, size=src.shape[2:], mode="bilinear", align_corners=False)
            y = up(torch.cat([y, src], 1))
        y = F.interpolate(y, size=s[4].shape[2:], mode="bilinear", align_corners=False)
        return torch.cat([y, s[4]], 1)[0].clone()          # 192 x H/2 x W/2


def main():
    torch.manual_seed(0)
    craft = E.CRAFT()
    craft.load_state_dict(weights.to_torch_state(weights.random_craft_state(SEED)))
    craft.eval()
    pages = []
    for i in range(5):
        pages.append(synth.title_page(9000 + i, 960, 704, True))
        pages.append(synth.book_cover(9100 + i, 960, 704, True))
    pages.append(synth.title_page(9200, 1280, 960, True))
    pages.append(synth.book_cover(9201, 1280, 960, True))
    feats, tgts = [], []
    for i, (rgb, mask) in enumerate(pages):
        t, l = synth.score_maps_for(mask, np.random.default_rng(i))
        feats.append(cached_inputs(craft, rgb))
        tgts.append(torch.from_numpy(np.stack([t, l])))
        print("cached", i, feats[-1].shape, flush=True)
    params = list(craft.upconv4.parameters()) + list(craft.conv_cls.parameters())
    for p in craft.parameters():
        p.requires_grad_(False)
    for p in params:
        p.requires_grad_(True)
    # the deep (upsampled) half of upconv4's input carries page-level low-frequency content that does not transfer
    # between pages: zero its weights and keep them zero, so the fitted tail is a local function of relu2_2
    w0 = craft.upconv4.conv[0].weight
    with torch.no_grad():
        w0[:, :64] = 0
    mask = torch.ones_like(w0); mask[:, :64] = 0
    w0.register_hook(lambda g: g * mask)
    opt = torch.optim.Adam(params, lr=2e-3)
    hp, hm = synth.title_page(2001, 1280, 960, True)
    hold_f = cached_inputs(craft, hp)[None]
    hold_t = synth.score_maps_for(hm)[0]
    rng = np.random.default_rng(0)
    C = 192
    t0 = time.time()
    for step in range(900):
        xb, yb = [], []
        for _ in range(4):
            k = int(rng.integers(len(feats)))
            f, t = feats[k], tgts[k]
            y0 = int(rng.integers(0, f.shape[1] - C + 1)); x0 = int(rng.integers(0, f.shape[2] - C + 1))
            xb.append(f[:, y0:y0 + C, x0:x0 + C]); yb.append(t[:, y0:y0 + C, x0:x0 + C])
        xb, yb = torch.stack(xb), torch.stack(yb)
        out = craft.conv_cls(craft.upconv4(xb))
        w = 1.0 + 4.0 * (yb > 0.3).float()
        loss = (w * (out - yb) ** 2).mean()
        opt.zero_grad(); loss.backward(); opt.step()
        if step % 100 == 0:
            with torch.no_grad():
                ho = craft.conv_cls(craft.upconv4(hold_f))[0, 0].numpy()
            print(step, float(loss), f"{time.time() - t0:.0f}s", "holdout frac>0.4: pred %.3f target %.3f" %
                  ((ho > 0.4).mean(), (hold_t > 0.4).mean()), flush=True)
        if step == 650:
            for g in opt.param_groups:
                g["lr"] = 5e-4
    out = {"seed": np.int64(SEED)}
    for name, mod in (("upconv4", craft.upconv4), ("conv_cls", craft.conv_cls)):
        for k, v in mod.state_dict().items():
            if k.endswith("num_batches_tracked"):
                continue
            out[f"{name}.{k}"] = v.detach().numpy().astype(np.float16)
    np.savez_compressed(os.path.join(os.path.dirname(weights.__file__), "data", "craft_probe.npz"), **out)
    print("saved", sum(v.size for v in out.values()))


if __name__ == "__main__":
    main()
