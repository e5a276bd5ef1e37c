#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of the detector's convolution launches (one page) into the text table and the
traffic JSON that profiles/ keeps and bench.py reads (roofline.traffic).

    ncu -i gpurun_out/r2_x3_conv.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_conv_summary.py raw.csv profiles/r2_x3_conv_ncu_full.summary.txt profiles/r2_conv_traffic.json "<title line>"
"""
import csv
import json
import re
import sys

LAYERS = [("conv1_1 (k_conv_stem: K = 32 stem gathered in shared memory, TMA-store epilogue) 32->64 @1/1", 32, 64, 1, 1.0), ("conv1_2 64->64 @1/1 (+pool)", 64, 64, 9, 1.0),
          ("conv2_1 64->128 @1/2", 64, 128, 9, 0.25), ("conv2_2 128->128 @1/2 (+pool, skip)", 128, 128, 9, 0.25),
          ("conv3_1 128->256 @1/4", 128, 256, 9, 1 / 16), ("conv3_2 256->256 @1/4", 256, 256, 9, 1 / 16),
          ("conv3_3 256->256 @1/4 (+pool)", 256, 256, 9, 1 / 16), ("conv4_1 256->512 @1/8", 256, 512, 9, 1 / 64),
          ("conv4_2 512->512 @1/8", 512, 512, 9, 1 / 64), ("conv4_3 512->512 @1/8 (+pool)", 512, 512, 9, 1 / 64),
          ("conv5_1 512->512 @1/16", 512, 512, 9, 1 / 256), ("conv5_2 512->512 @1/16", 512, 512, 9, 1 / 256),
          ("fc6 512->1024 dil6 @1/16", 512, 1024, 9, 1 / 256), ("fc7 1x1 1024->1024", 1024, 1024, 1, 1 / 256),
          ("up1a 1x1 1536->512", 1536, 512, 1, 1 / 256), ("up1b 512->256", 512, 256, 9, 1 / 256),
          ("up2a 1x1 768->256 @1/8", 768, 256, 1, 1 / 64), ("up2b 256->128", 256, 128, 9, 1 / 64),
          ("up3a 1x1 384->128 @1/4", 384, 128, 1, 1 / 16), ("up3b 128->64", 128, 64, 9, 1 / 16),
          ("up4a 1x1 192->64 @1/2", 192, 64, 1, 0.25), ("up4b 64->32", 64, 32, 9, 0.25), ("cls0 32->32", 32, 32, 9, 0.25),
          ("cls1 32->32", 32, 32, 9, 0.25), ("cls2 32->16", 32, 16, 9, 0.25)]


def main():
    raw, out_txt, out_json = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else ""
    px = float(sys.argv[5]) if len(sys.argv) > 5 else 1920 * 1440
    mma = int(sys.argv[6]) if len(sys.argv) > 6 else 3
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    g = lambda r, k: float(r[col[k]]) if r[col[k]] not in ("", "n/a") else 0.0      # noqa: E731
    bscale = {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}
    tscale = {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
    lines = [title, "layer | kernel | grid | time us | DRAM read MB | DRAM write MB | DRAM %peak | tensor pipe active % | L2 throughput % | "
             "L1/smem throughput % | regs | dyn smem KB | algorithmic GFLOP | algorithmic TFLOP/s | MMA TFLOP/s"]
    tot_t = tot_r = tot_w = tot_f = 0.0
    n = 0
    for r, (name, cin, cout, taps, frac) in zip(data, LAYERS):
        t = g(r, "gpu__time_duration.sum") * tscale.get(units[col["gpu__time_duration.sum"]], 1.0)
        rd = g(r, "dram__bytes_read.sum") * bscale.get(units[col["dram__bytes_read.sum"]], 1.0)
        wr = g(r, "dram__bytes_write.sum") * bscale.get(units[col["dram__bytes_write.sum"]], 1.0)
        gf = 2.0 * px * frac * cin * cout * taps / 1e9
        kern = re.sub(r"\(CUtensorMap.*", "", r[col["Kernel Name"]]).replace("void bbocr::<unnamed>::", "")
        lines.append(f"{name} | {kern} | {r[col['Grid Size']]} | {t:.1f} | {rd:.1f} | {wr:.1f} | "
                     f"{g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                     f"{g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                     f"{g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                     f"{g(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {r[col['launch__registers_per_thread']]} | "
                     f"{g(r, 'launch__shared_mem_per_block_dynamic'):.1f} | {gf:.1f} | {gf / t * 1e3:.0f} | {mma * gf / t * 1e3:.0f}")
        tot_t += t; tot_r += rd; tot_w += wr; tot_f += gf; n += 1
    lines.append(f"TOTAL {n} launches: {tot_t:.1f} us, DRAM read {tot_r:.1f} MB + write {tot_w:.1f} MB = {tot_r + tot_w:.1f} MB per page; per launch "
                 f"{(tot_r + tot_w) / n:.2f} MB; algorithmic {tot_f:.0f} GFLOP -> {tot_f / tot_t * 1e3:.0f} TFLOP/s algorithmic, {mma * tot_f / tot_t * 1e3:.0f} TFLOP/s of "
                 f"bf16 MMA work over the serialised cold-cache launches")
    open(out_txt, "w").write("\n".join(lines) + "\n")
    json.dump({"dram_bytes_per_launch": (tot_r + tot_w) * 1e6 / n, "dram_bytes_per_page": (tot_r + tot_w) * 1e6, "launches": n,
               "source": out_txt, "note": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, one 1920x1440 page"}, open(out_json, "w"))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
