"""Dump bf16x3 CRAFT score maps of a fixed page set into an .npz (argv[1]).

tests/test_gpu_detector.py::test_fused_stem_equals_two_kernel_path runs this twice in sub-processes, with and without
BBOCR_STEM_FUSED=0 (the switch is read once per process), and compares the files bitwise: k_conv_stem (gather + conv1_1 in one
kernel) must reproduce k_im2col_rgb_split + k_conv_tc exactly.  With "time" as argv[2] it also prints the score_maps time of the
1920x1440 page (host-timed, after warm-up) for a quick same-box A/B."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PAGES = [("title", 31, 640, 480), ("cover", 32, 333, 250), ("title", 33, 1920, 1440), ("cover", 34, 96, 64)]


def main():
    import bbocr_b200
    from bbocr_b200 import synth, weights
    reader = bbocr_b200.Reader(["en"], gpu=True, verbose=False, precision="bf16x3", craft_state=weights.calibrated_craft_state(),
                               crnn_state=weights.calibrated_crnn_state())
    out = {}
    for i, (kind, seed, w, h) in enumerate(PAGES):
        img = (synth.title_page if kind == "title" else synth.book_cover)(seed, w, h)
        t, l, _ = reader.score_maps(img)
        out[f"t{i}"], out[f"l{i}"] = t, l
    # the canvas-resize branch (INTER_LINEAR u8 resize in front of the stem, image smaller than the 32-aligned canvas)
    out["t_resized"], out["l_resized"], _ = reader.score_maps(synth.title_page(35, 700, 500), canvas_size=480)
    # a batch of pages of one geometry through the batched detector (NIMG > 1 inside one launch)
    pages = [synth.title_page(40 + k, 640, 480) for k in range(3)]
    res = reader.readtext_batched(pages)
    out["batched"] = np.array([repr([(np.asarray(b).tolist(), s) for b, s, _ in r]) for r in res])
    np.savez(sys.argv[1], **out)
    if len(sys.argv) > 2 and sys.argv[2] == "time":
        img = synth.title_page(33, 1920, 1440)
        for _ in range(3):
            reader.score_maps(img)
        t0 = time.perf_counter()
        for _ in range(20):
            reader.score_maps(img)
        print("score_maps 1920x1440 ms", (time.perf_counter() - t0) / 20 * 1e3, "fused",
              os.environ.get("BBOCR_STEM_FUSED", "1"))


if __name__ == "__main__":
    main()
