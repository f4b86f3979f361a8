"""Per-kernel shares of ONE training step from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python bench.py --steps 2 --warmup 1 --no-graph ...`).
A step is cut at the optimizer: from the launch after one k_adam group to the end of the next.
usage: python tools/launch_shares.py gpurun_out/launches_<tag>.csv profiles/<tag>_launch_shares.csv [profiles/<tag>_launches_raw.csv]"""
import collections, csv, re, sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = []
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        ns = v * 1000 if u in ("us", "usecond") else (v * 1e6 if u.startswith("ms") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name).replace("mapnet::", "")
        rows.append((name, ns))
    adam = [i for i, (n, _) in enumerate(rows) if n.startswith("k_adam")]
    # k_adam runs in groups (one launch per parameter run, each followed by its step counter): group boundaries
    groups = [adam[0]]
    for a, b in zip(adam, adam[1:]):
        if b - a > 4:
            groups.append(b)
    ends = []
    for g in groups:
        e = g
        while e + 1 < len(rows) and (rows[e + 1][0].startswith("k_adam") or rows[e + 1][0].startswith("k_inc_i32")):
            e += 1
        ends.append(e)
    a, b = ends[-2] + 1, ends[-1] + 1
    step = rows[a:b]
    agg = collections.OrderedDict()
    for n, ns in step:
        d = agg.setdefault(n, [0, 0.0]); d[0] += 1; d[1] += ns
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# %s\n" % " ".join(sys.argv))
        f.write("# exactly ONE eager training step (after one optimizer group .. end of the next): %d launches, %.1f us "
                "serialised, cold-cache: compare SHARES, not absolutes\n" % (len(step), tot / 1e3))
        f.write("kernel,launches,total_us,share_pct,avg_us\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.1f,%.2f,%.1f\n' % (k, v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / 1e3 / v[0]))
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            f.write("index,kernel,dur_ns\n")
            for i, (n, ns) in enumerate(step):
                f.write('%d,"%s",%.0f\n' % (i, n, ns))
    conv = sum(v[1] for k, v in agg.items() if k.startswith("k_tc_"))
    print("%d launches, %.1f us, conv share %.3f" % (len(step), tot / 1e3, conv / tot))


if __name__ == "__main__":
    main()
