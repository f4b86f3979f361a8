"""Summarise an `ncu --set full` report into profiles/ (run HERE, on the report gpurun brought back).

    python tools/ncu_summary.py gpurun_out/conv_full.ncu-rep r02a
    python tools/ncu_summary.py --csv gpurun_out/conv_full_fwd_raw.csv gpurun_out/conv_full_bwd_raw.csv r02a

writes profiles/<tag>_ncu_full_summary.csv (one row per captured launch: duration, tensor-pipe activity, DRAM
bytes and rate, L2 throughput, registers) and profiles/ncu_traffic.json (dram read + write bytes per launch of the
dominant conv kernel = the launch list's biggest total among the captured kernels), which bench.py reports as
`roofline.traffic`.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_DIR = os.environ.get("NCU_SUMMARY_OUT", os.path.join(ROOT, "profiles"))      # tests write elsewhere

WANT = [  # (column label, exact ncu metric name)
    ("dur_ns", "gpu__time_duration.sum"),
    ("tensor_pipe_pct_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_pipe_pct_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("utcmma_insts", "smsp__sass_inst_executed_op_utcmma.sum"),
    ("tma_ld_bytes", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"),
    ("dram_read_bytes", "dram__bytes_read.sum"),
    ("dram_write_bytes", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("smem_dyn", "launch__shared_mem_per_block_dynamic"),
    ("grid", "launch__grid_size"),
]

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}


def parse(raw, out_rows, per_kernel, first_hdr):
    """one `--page raw --csv` export (its own header + units rows) -> records"""
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[hdr_i], rows[hdr_i + 1], rows[hdr_i + 2:]
    if not first_hdr:
        first_hdr.extend(hdr)
    col = {label: hdr.index(metric) for label, metric in WANT if metric in hdr}
    kn = hdr.index("Kernel Name")
    for r in data:
        if len(r) <= kn:
            continue
        name = re.sub(r"^(void )?(mapnet::)?", "", r[kn])
        name = re.sub(r"\(.*$", "", name)
        rec = {"kernel": name}
        for label, j in col.items():
            try:
                v = float(r[j].replace(",", ""))
            except ValueError:
                v = None
            if v is not None and label in ("dur_ns", "dram_read_bytes", "dram_write_bytes", "tma_ld_bytes"):
                v *= UNIT.get(units[j], 1.0)
            rec[label] = v
        out_rows.append(rec)
        k = per_kernel.setdefault(name, {"n": 0, "dur": 0.0, "bytes": 0.0})
        k["n"] += 1
        k["dur"] += rec.get("dur_ns") or 0.0
        k["bytes"] += (rec.get("dram_read_bytes") or 0.0) + (rec.get("dram_write_bytes") or 0.0)
    return col


def main():
    args = sys.argv[1:]
    out_rows, per_kernel, hdr, col = [], {}, [], {}
    if args and args[0] == "--csv":          # one or more `ncu -i x.ncu-rep --page raw --csv` exports made on the GPU box
        tag = args[-1]
        for f in args[1:-1]:
            txt = open(f).read()
            if "Kernel Name" in txt:
                col.update(parse(txt, out_rows, per_kernel, hdr))
    else:
        rep, tag = args[0], args[1]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        col.update(parse(raw, out_rows, per_kernel, hdr))
    labels = ["kernel"] + [l for l, _ in WANT if l in col]
    path = os.path.join(OUT_DIR, "%s_ncu_full_summary.csv" % tag)
    with open(path, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; one row per captured launch; metric columns: "
                + "; ".join("%s=%s" % (l, dict(WANT)[l]) for l in labels[1:]) + "; byte / time columns in bytes / ns\n")
        w = csv.writer(f)
        w.writerow(labels)
        for rec in out_rows:
            w.writerow([rec.get(l) for l in labels])
    print("wrote", path, len(out_rows), "launches")
    conv = {k: v for k, v in per_kernel.items() if k.startswith(os.environ.get("NCU_SUMMARY_PREFIX", "k_tc_"))}
    if conv:
        top = max(conv, key=lambda k: conv[k]["dur"])
        t = conv[top]
        traffic = {"kernel": top, "launches_captured": t["n"], "dram_bytes_per_launch": t["bytes"] / t["n"],
                   "avg_duration_us": t["dur"] / t["n"] / 1e3, "source": os.path.basename(path)}
        dig = os.path.join(ROOT, "gpurun_out", "ncu_sources_digest.txt")
        if os.path.exists(dig):          # digest of the kernel sources the capture ran (bench.py refuses a stale capture)
            traffic["sources_digest"] = open(dig).read().strip()
        with open(os.path.join(OUT_DIR, "ncu_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
        print("ncu_traffic.json:", traffic)


if __name__ == "__main__":
    main()
