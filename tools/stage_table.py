"""Per-stage view of the conv launches of ONE training step, from the committed ncu launch list.

    python tools/stage_table.py profiles/r02c_launches_raw.csv [peak_tflops] > profiles/r02c_conv_by_stage.txt

The launch list (tools/launch_shares.py: `ncu --metrics gpu__time_duration.sum --clock-control none`, one eager step of
posenet_bs64) is in launch order, which is the trunk's schedule (geomapnet_b200/csrc/net.cu): forward = stem, then per
BasicBlock conv1, conv2[, downsample]; backward = blocks in reverse, per block dgrad(conv2) then dgrad(conv1) (a
downsampling block's conv1 dgrad carries the 1x1 shortcut dgrad in the same launch), wgrad(conv2), wgrad(conv1)
[, wgrad(downsample)], the stem's wgrad last.  Every conv launch is matched to its layer by position and its ALGORITHMIC
FLOPs (B = 64, 256x256) are divided by its duration.  Durations under ncu are serialised and cold-cache: compare stages
with each other and against bench.py's event-bracketed per-class numbers, not as absolute step time."""
import csv
import sys

B, HW = 64, 64          # frames, feature-map side after the stem pool (256x256 input)
STAGES = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]


def layers():
    stem = 2 * B * 128 * 128 * 64 * 147
    blocks, inpl, h = [], 64, HW
    for li, (pl, nb, st) in enumerate(STAGES):
        for b in range(nb):
            s = st if b == 0 else 1
            ho = h // s
            f1 = 2 * B * ho * ho * pl * 9 * inpl
            f2 = 2 * B * ho * ho * pl * 9 * pl
            fd = 2 * B * ho * ho * pl * inpl if (s != 1 or inpl != pl) else 0
            blocks.append(("layer%d" % (li + 1), f1, f2, fd))
            inpl, h = pl, ho
    return stem, blocks


def main():
    path = sys.argv[1]
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else 1653.4      # burst bf16 TFLOP/s measured in the capture's session
    rows = list(csv.reader(open(path)))[1:]
    L = [(r[1], float(r[2]) / 1e3) for r in rows]                   # (kernel, us)
    stem, blocks = layers()
    split = next(i for i, (k, _) in enumerate(L) if k.startswith("k_gap_bwd"))
    fw = [x for x in L[:split] if x[0].startswith("k_tc_conv")]
    dg = [x for x in L[split:] if x[0].startswith("k_tc_conv")]
    wg = [x for x in L[split:] if x[0].startswith("k_tc_wgrad")]
    f_order = [("stem", stem)]
    for st, f1, f2, fd in blocks:
        f_order += [(st, f1), (st, f2)] + ([(st, fd)] if fd else [])
    d_order, w_order = [], []
    for st, f1, f2, fd in reversed(blocks):
        d_order += [(st, f2), (st, f1 + fd)]
        w_order += [(st, f2), (st, f1)] + ([(st, fd)] if fd else [])
    w_order.append(("stem", stem))
    assert len(fw) == len(f_order) == 36 and len(dg) == len(d_order) == 32 and len(wg) == len(w_order) == 36, \
        (len(fw), len(dg), len(wg))
    print("# %s: conv launches of one eager posenet_bs64 step by ResNet stage; peak = %.1f TFLOP/s (burst)" % (path, peak))
    grand_us = grand_fl = 0.0
    for name, lst, order in (("fprop", fw, f_order), ("dgrad", dg, d_order), ("wgrad", wg, w_order)):
        agg = {}
        for (k, us), (st, fl) in zip(lst, order):
            a = agg.setdefault(st, [0.0, 0.0, 0, set()])
            a[0] += us; a[1] += fl; a[2] += 1; a[3].add(k.split("(")[0])
        tu = sum(a[0] for a in agg.values()); tf = sum(a[1] for a in agg.values())
        grand_us += tu; grand_fl += tf
        print("\n%s | launches | us | GFLOP | TFLOP/s | of peak | share of %s time | share of FLOPs | kernels" % (name, name))
        for st in ("stem", "layer1", "layer2", "layer3", "layer4"):
            if st not in agg:
                continue
            us, fl, n, ks = agg[st]
            print("%-6s | %2d | %6.1f | %6.1f | %4.0f | %.2f | %4.1f %% | %4.1f %% | %s"
                  % (st, n, us, fl / 1e9, fl / us / 1e6, fl / us / 1e6 / peak, 100 * us / tu, 100 * fl / tf, ", ".join(sorted(ks))))
        print("%-6s | %2d | %6.1f | %6.1f | %4.0f | %.2f" % ("all", sum(a[2] for a in agg.values()), tu, tf / 1e9,
                                                          tf / tu / 1e6, tf / tu / 1e6 / peak))
    print("\nall conv launches: %.1f us, %.1f GFLOP, %.0f TFLOP/s = %.2f of peak"
          % (grand_us, grand_fl / 1e9, grand_fl / grand_us / 1e6, grand_fl / grand_us / 1e6 / peak))


if __name__ == "__main__":
    main()
