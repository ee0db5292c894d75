"""Achieved HBM bandwidth of the step's elementwise / reduction kernels against the measured B200 copy bandwidth.

    python tools/hbm_roofline.py [launches.csv] [epilogue.csv] > profiles/rNN_hbm_kernels_roofline.txt     (CPU only)

Input: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch lists of eager C3 train steps (B = 256,
T_a = 250, T_v = 64, d = 768, d_ff = 2048, bf16).  Shapes are not in the launch list: they follow from the ORDER of the
launches inside a step, which is fixed by the model (forward: audio encoder, video encoder, fusion encoder, adaptors,
heads; backward: the reverse, video before audio) — the decoding below is checked against the duration ratios.
achieved = ALGORITHMIC bytes per launch (every tensor the kernel must read or write once, bf16 = 2 B)
/ the median duration of that launch position; peak = MEASURED_PEAKS.json hbm_gbs (torch copy, read + write)."""
import collections
import csv
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, TA, TV, D, DFF = 256, 250, 64, 768, 2048
ROWS = {"audio": B * TA, "video": B * TV, "fusion": B * (TA + TV)}


def load(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
    out = []
    for r in rows:
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        out.append((name, r[gi], float(r[vi]) / 1e3))
    return out


def steps_of(launches, marker="adam_kernel"):
    """Split into steps at the optimizer launch; keep only steps with the modal number of launches."""
    steps, cur = [], []
    for l in launches:
        cur.append(l)
        if l[0].startswith(marker):
            steps.append(cur)
            cur = []
    n = collections.Counter(len(s) for s in steps).most_common(1)[0][0]
    return [s for s in steps if len(s) == n]


def main():
    launches = load(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_step_launches_ncu.csv"))
    epi_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r01_bwd_epilogue_launches_ncu.csv")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    steps = steps_of(launches)
    table = []      # (kernel, what, bytes, [durations us])

    def by_position(prefix, labels, bytes_of, source=None):
        per = collections.defaultdict(list)
        for s in (source or steps):
            ls = [l for l in s if l[0].startswith(prefix)]
            if len(ls) != len(labels):
                continue
            for lab, l in zip(labels, ls):
                per[lab].append(l[2])
        for lab in dict.fromkeys(labels):
            if per[lab]:
                table.append((prefix, lab, bytes_of(lab), per[lab]))

    enc = lambda lab: ROWS[lab.split()[0]]
    # LayerNorm: forward reads x, writes y (+ 8 B of statistics per row); backward reads dy and x, writes dx
    by_position("layernorm_fwd_kernel", ["audio"] * 3 + ["video"] * 3 + ["fusion"] * 3, lambda l: enc(l) * (2 * D * 2 + 8))
    by_position("layernorm_bwd_kernel", ["fusion"] * 3 + ["video"] * 3 + ["audio"] * 3, lambda l: enc(l) * (3 * D * 2 + 8))
    # kernels whose grid identifies the shape (one block per 2048 elements): 24000 blocks = the audio tensor
    # (256 x 250 x 768), 6144 = the video tensor after the embedding (256 x 64 x 768), 4096 = the raw video input (x 512)
    def by_grid(prefix, grids, bytes_of):
        per = collections.defaultdict(list)
        for s in steps:
            for l in s:
                if l[0].startswith(prefix) and l[1] in grids:
                    per[grids[l[1]]].append(l[2])
        for lab, durs in per.items():
            table.append((prefix, lab, bytes_of(lab), durs))

    AUD, VID, VIN = "(24000, 1, 1)", "(6144, 1, 1)", "(4096, 1, 1)"
    # concat along T (fusion input): forward copies each modality into the fused buffer, backward copies the slices out
    by_grid("concat_rows_kernel", {AUD: "audio", VID: "video"}, lambda l: enc(l) * D * 2 * 2)
    by_position("rowzero_kernel", ["fusion"], lambda l: enc(l) * D * 2)
    by_grid("meanpool_bwd_kernel", {AUD: "audio", VID: "video"}, lambda l: enc(l) * D * 2)      # writes dx (reads B x D)
    by_grid("cast_kernel<float, __nv_bfloat16>", {AUD: "audio input fp32->bf16 (L2-warm after the H2D copy)",
                                                   VIN: "video input fp32->bf16 (L2-warm after the H2D copy)"},
            lambda l: (B * TA * D if l[0] == "a" else B * TV * 512) * 6)
    by_position("adam_kernel", ["19.7 M parameters"], lambda l: 19_697_668 * 28)
    # dropout / ReLU backward + bias gradient: its own launch list (taken after the kernel's rewrite), backward order
    if os.path.exists(epi_path):
        epi = load(epi_path)
        n = 19
        esteps = [epi[i:i + n] for i in range(0, len(epi) - n + 1, n)]
        labels = ["head", "head hidden", "head", "head hidden",
                  "video adaptor (ReLU+dropout, N=768)", "audio adaptor (ReLU+dropout, N=768)",
                  "fusion linear2 (dropout, N=768)", "fusion linear1 (ReLU+dropout, N=2048)", "fusion out_proj (dropout, N=768)",
                  "fusion in_proj (bias only, N=2304)",
                  "video linear2 (dropout, N=768)", "video linear1 (ReLU+dropout, N=2048)", "video out_proj (dropout, N=768)",
                  "video in_proj (bias only, N=2304)", "video embedding (ReLU, N=768)",
                  "audio linear2 (dropout, N=768)", "audio linear1 (ReLU+dropout, N=2048)", "audio out_proj (dropout, N=768)",
                  "audio in_proj (bias only, N=2304)"]

        def epi_bytes(lab):
            if lab.startswith("head"):
                return 0
            rows = enc(lab)
            width = int(lab.split("N=")[1].rstrip(")"))
            passes = 1 if "bias only" in lab else (3 if "ReLU" in lab else 2)     # dout [+ out for the ReLU mask] [+ dz]
            return rows * width * 2 * passes
        by_position("bwd_epilogue_kernel", labels, epi_bytes, source=esteps)

    print(f"# HBM-bound kernels of the C3 train step (B=256, bf16): achieved = algorithmic bytes / ncu duration (cold cache,")
    print(f"# serialised launches); peak = {peak:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs, torch copy read+write)")
    print(f"# {'kernel':34s} {'launch':40s} {'MB':>8s} {'us':>7s} {'GB/s':>7s} {'frac':>5s}  n")
    for k, lab, nbytes, durs in table:
        if not nbytes:
            continue
        us = statistics.median(durs)
        gbs = nbytes / us / 1e3
        print(f"  {k[:34]:34s} {lab[:40]:40s} {nbytes / 1e6:8.1f} {us:7.1f} {gbs:7.0f} {gbs / peak:5.2f}  {len(durs)}")


if __name__ == "__main__":
    main()
