#!/usr/bin/env python
"""bench.py -- images/sec of the MapNet/PoseNet ResNet-34 training step on B200.

    python bench.py --gpus N --steps K --warmup W [--workload posenet_bs64|mapnet_n32t3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

A "step" is common/train.py:339-361: forward, criterion, backward, [one NCCL
allreduce of the flat gradient buffer], Adam.  `value` = whole-job images/sec
with inputs resident in HBM; `e2e` = the same step through the reference-facing
nn.Module surface with HOST (pinned) inputs, H2D copies and loss.item() inside
the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: PoseNet ResNet-34 bs64 256x256, PoseNetCriterion
    "posenet_bs64": dict(kind="posenet", N=64, T=1, H=256, W=256, lr=1e-4, wd=5e-4, clip=0.0),
    # BASELINE.json configs[2]/[3]: MapNet steps=3 bs32 (96 frames), MapNetCriterion
    "mapnet_n32t3": dict(kind="mapnet", N=32, T=3, H=256, W=256, lr=1e-4, wd=5e-4, clip=0.0),
    # BASELINE.json configs[4]: MapNet++ steps=5 bs16 through MFOnline (160 frames), MapNetOnlineCriterion
    "mapnetpp_n16t10": dict(kind="online", N=16, T=10, H=256, W=256, lr=1e-5, wd=0.0, clip=5.0),
}


def usable_cores():
    """Host threads this process can really use: CPU affinity capped by the cgroup quota
    (os.cpu_count() over-reports inside containers and oversubscription wrecks the CPU arm)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    cap = os.environ.get("BENCH_CPU_THREADS")
    if cap:
        n = max(1, int(cap))
    return n


def tune_cpu_threads(kind):
    """Pick the torch thread count that makes the CPU arm FASTEST on this host (all usable
    cores is not always best: NUMA / SMT oversubscription).  A 4-frame step per candidate."""
    import torch
    from oracle import weights, mapnet_oracle as O
    n = usable_cores()
    cands = sorted({c for c in (n, n // 2, 64, 32, 16, 8) if 1 <= c <= n}, reverse=True)
    st = weights.make_state(7)
    cfg = dict(kind="posenet", N=4, T=1, H=256, W=256)
    x, targ = weights.make_inputs(cfg, 5)
    sv = dict(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0)
    best, best_t = n, None
    for c in cands:
        torch.set_num_threads(c)
        O.train_step("posenet", st, x, targ, sv)             # warm
        t0 = time.time()
        O.train_step("posenet", st, x, targ, sv)
        dt = time.time() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tc_burst=float(d["bf16_tflops"]),
                    tc_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant conv kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written from the CSV export by tools/ncu_summary.py).
    The capture records the digest of the kernel sources it was taken with; a capture of OTHER sources than the ones
    this run was built from is reported as stale (traffic = None) instead of being passed off as a measurement of
    this build.  None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None
    try:
        from geomapnet_b200 import build as _b
        d["sources_digest_now"] = _b._digest()[:16]
        d["stale"] = bool(d.get("sources_digest")) and d["sources_digest"] != d["sources_digest_now"]
    except Exception:
        d["stale"] = None
    return d


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop.is_set():
                    break
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        if self.proc is not None:
            try:
                self.proc.kill()
            except Exception:
                pass

    def summary(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, c in self.rows:
            if ts < t0 or ts > t1 + 0.2 or len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx = max(mx, float(c[1]))
            except ValueError:
                continue
            for i, nm in enumerate(names):
                if c[3 + i].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            for ts, c in self.rows[-3:]:
                try:
                    sm.append(float(c[0])); mx = max(mx, float(c[1]))
                except Exception:
                    pass
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(mx or None),
                    reasons=sorted(reasons), samples=len(sm))


def make_batches(cfg, n, seed, device=None, pinned=False):
    """Synthetic inputs of the workload's shape; distinct per step."""
    import torch
    from oracle import weights
    out = []
    for i in range(n):
        x, targ = weights.make_inputs(dict(kind=cfg["kind"], N=cfg["N"], T=cfg["T"], H=cfg["H"], W=cfg["W"]), seed + i)
        if device is not None:
            x, targ = x.to(device), targ.to(device)
        elif pinned:
            x, targ = x.pin_memory(), targ.pin_memory()
        out.append((x, targ))
    return out


def frames(cfg):
    return cfg["N"] * cfg["T"]


# --------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation (oracle port) on the host cores
# --------------------------------------------------------------------------------
def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    import torch
    from oracle import weights, mapnet_oracle as O
    cores = tune_cpu_threads(cfg["kind"])
    st = weights.make_state(7)
    sv = dict(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0)
    # bounded sample: each timed step is one full step of the workload at a reduced
    # tuple count so that K+W steps end within minutes on any host
    ref_frames = args.ref_frames if args.ref_frames else frames(cfg)       # default: the FULL workload (same config)
    n_tuples = min(cfg["N"], max(1, ref_frames // cfg["T"]))
    scfg = dict(cfg, N=n_tuples)
    batches = make_batches(scfg, args.steps + args.warmup, 100)
    times = []
    tr = O.OracleTrainer(cfg["kind"], st, sv, lr=cfg["lr"], weight_decay=cfg["wd"], max_grad_norm=cfg["clip"],
                         droprate=args.droprate)
    for i, (x, targ) in enumerate(batches):
        t0 = time.time()
        tr.step(x, targ)
        dt = time.time() - t0
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    val = frames(scfg) / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": "images/sec", "value": val, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "frames_per_step": frames(scfg),
                   "note": "oracle port of the reference CPU path (torch CPU ops); thread count = fastest of "
                           "{all usable cores, half, 64, 32, 16, 8} on this host, %d usable" % usable_cores()},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d-frame steps of %s (%s)"
                                   % (frames(scfg), args.workload, "the full workload" if frames(scfg) == frames(cfg)
                                      else "bounded sample of the %d-frame workload" % frames(cfg))},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------
def run_b200(args, cfg, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import torchvision
    from geomapnet_b200 import _lib
    from geomapnet_b200.models.posenet import PoseNet, MapNet
    from geomapnet_b200.common.criterion import PoseNetCriterion, MapNetCriterion, MapNetOnlineCriterion
    from geomapnet_b200.common.optimizer import Optimizer
    from geomapnet_b200.ddp import FlatDataParallel
    from oracle import mapnet_oracle as O      # FLOP model only (roofline denominator)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(7)
    fe = torchvision.models.resnet34(weights=None)
    net = PoseNet(fe, droprate=args.droprate, pretrained=False, filter_nans=(cfg["kind"] == "online"),
                  precision=args.precision, seed=7 + rank)
    model = net if cfg["kind"] == "posenet" else MapNet(net)
    def make_crit():
        kw = dict(sax=0.0, saq=-3.0)
        if cfg["kind"] == "posenet":
            return PoseNetCriterion(learn_beta=True, **kw)
        if cfg["kind"] == "mapnet":
            return MapNetCriterion(srx=0.0, srq=-3.0, learn_beta=True, learn_gamma=True, **kw)
        return MapNetOnlineCriterion(srx=0.0, srq=-3.0, learn_beta=True, learn_gamma=True, **kw)

    crit = make_crit()
    params = [{"params": model.parameters()}, {"params": list(crit.parameters())}]
    opt = Optimizer(params=params, method="adam", base_lr=cfg["lr"], weight_decay=cfg["wd"])
    model.cuda(); crit.cuda(); model.train()
    dp = None

    def eager_step(x, targ):
        out = model(x)
        loss = crit(out, targ)
        opt.learner.zero_grad()
        loss.backward()
        scale = 1.0
        if dp is not None:
            scale = dp.allreduce_grads()
        opt.learner.step(grad_scale=scale, max_grad_norm=cfg["clip"])
        return loss

    nb = max(2, min(4, args.steps))
    dev_batches = make_batches(cfg, nb, 1000 * (rank + 1), device=dev)
    # first step: builds the arena / flattens parameters
    eager_step(*dev_batches[0])
    if world > 1:
        dp = FlatDataParallel(model, crit, overlap=(os.environ.get("MAPNET_DDP_OVERLAP", "0") == "1"))
        dp.broadcast_parameters()
    L = _lib.lib()
    step = eager_step
    graph_launches = None
    if args.graph:
        # the same modules, captured once as two CUDA graphs and replayed (geomapnet_b200/graph.py)
        from geomapnet_b200.graph import GraphedTrainStep
        gstep = GraphedTrainStep(model, crit, opt, dev_batches[0][0], dev_batches[0][1], dp=dp,
                                 max_grad_norm=cfg["clip"], warmup=2)
        graph_launches = gstep.kernels_per_step
        step = gstep

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: inputs resident in HBM ----------------
    for i in range(args.warmup):
        step(*dev_batches[i % nb])
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    lc0 = L.mapnet_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(*dev_batches[i % nb])
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = (L.mapnet_launch_count() - lc0) // max(1, args.steps)
    if args.graph:
        launches = graph_launches       # kernels inside the replayed graphs (counted while capturing)
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * frames(cfg) / (ms_step / 1000.0)
    clocks = sampler.summary(t_wall0, t_wall1) if sampler else None

    # ---------------- e2e: host (pinned) inputs, H2D + loss.item() inside the timed region ----------
    host_batches = make_batches(cfg, nb, 2000 * (rank + 1), pinned=True)
    xd = torch.empty_like(dev_batches[0][0]); td = torch.empty_like(dev_batches[0][1])

    # Inputs start in pinned HOST memory every step (common/train.py:341,347 `.cuda(async=True)` of a
    # pin_memory DataLoader batch).  The H2D copy of batch i+1 runs on a copy stream while step i
    # computes (what pinned memory + async copy exist for); every step still ends with loss.item().
    copy_stream = torch.cuda.Stream(dev)
    stage_x = [torch.empty_like(dev_batches[0][0]) for _ in range(2)]
    stage_t = [torch.empty_like(dev_batches[0][1]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        xh, th = host_batches[i % nb]
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            stage_x[b].copy_(xh, non_blocking=True)
            stage_t[b].copy_(th, non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_loop(n):
        cur = torch.cuda.current_stream(dev)
        for b in range(2):
            consumed[b].record(cur)
        prefetch(0)
        last = None
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            b = i % 2
            cur.wait_event(ready[b])
            loss = step(stage_x[b], stage_t[b])
            consumed[b].record(cur)
            last = loss.item()                  # common/train.py:361  D2H + sync every step
        return last

    e2e_loop(min(2, args.warmup))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(args.steps)
    f1.record()
    barrier()
    e2e_ms = max_over_ranks(f0.elapsed_time(f1)) / args.steps
    e2e_val = world * frames(cfg) / (e2e_ms / 1000.0)

    # eager (no CUDA graph) step through the plain nn.Module calls, for transparency
    eager_ms = None
    if args.graph:
        net._graph_rng = False
        for i in range(3):
            eager_step(*dev_batches[i % nb])
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for i in range(args.steps):
            eager_step(*dev_batches[i % nb])
        h1.record()
        barrier()
        eager_ms = max_over_ranks(h0.elapsed_time(h1)) / args.steps
        step = eager_step
    h2d = host_batches[0][0].numel() * 4 + host_batches[0][1].numel() * 4
    if sampler:
        sampler.stop()

    # ---------------- roofline of the dominant kernel class (conv engines), rank 0 ----------------
    roof = None
    peaks = _peaks()
    import ctypes
    trunk = net._trunks[(dev.index, cfg["H"], cfg["W"])]
    psteps = 3
    # The profiled steps run eagerly with a CUDA-event pair around every conv launch.  Two unprofiled
    # steps are enqueued first WITHOUT a sync, so the device is busy while the host runs ahead and the
    # profiled launches are already queued when the device reaches them: no host-side gap lands
    # inside a bracket.  What remains is the isolation itself: a bracketed launch cannot overlap its prologue /
    # pipeline fill / tail with its neighbours as it does in the CUDA-graph step, +5-7 us per launch
    # (profiles/r01d_conv_microbench.txt, block 3 vs block 2) -- `achieved` is therefore a LOWER bound.
    for i in range(2):
        step(*dev_batches[i % nb])
    if rank == 0:
        _lib.check(L.mapnet_profile(trunk.h, 1), "mapnet_profile")
    for i in range(psteps):                 # every rank steps (the allreduce is a collective)
        step(*dev_batches[i % nb])
    barrier()
    if rank == 0:
        B = frames(cfg)
        ms3 = (ctypes.c_double * 3)(); fl3 = (ctypes.c_double * 3)(); n3 = (ctypes.c_int * 3)()
        _lib.check(L.mapnet_profile_read(trunk.h, ms3, fl3, n3), "mapnet_profile_read")
        _lib.check(L.mapnet_profile(trunk.h, 0), "mapnet_profile")
        tot_ms = sum(ms3); tot_fl = sum(fl3); tot_n = sum(n3)
        achieved = tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
        peak = peaks["tc_sustained"]
        per_class = {}
        for k, nm in enumerate(("fprop", "dgrad", "wgrad")):
            if ms3[k] > 0:
                per_class[nm] = {"tflops": fl3[k] / (ms3[k] * 1e-3) / 1e12, "ms_per_step": ms3[k] / psteps,
                                 "launches_per_step": n3[k] // psteps}
        peak = peaks["tc_burst"]      # every bracketed launch runs alone for 10-50 us at full clocks: the burst figure
        kname = {"bf16": "k_tc_conv / k_tc_conv2 / k_tc_conv_halo / k_tc_wgrad(2) (all conv launches of a step, bf16 operands)",
                 "tc_split": "k_tc_conv / k_tc_conv2 / k_tc_wgrad(2) (all conv launches of a step; fp16 hi/lo operand planes, "
                             "4 tcgen05 MMAs per algorithmic product: the tensor pipe executes 4x the algorithmic FLOPs)"}
        nt = _ncu_traffic()
        roof = {"bound": "tensor", "kernel": kname.get(args.precision, "k_conv_simt (fp32 CUDA-core engine)"),
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "%s (MEASURED_PEAKS.json bf16_tflops, burst)" % peaks["src"],
                "frac_of_sustained_peak": achieved / peaks["tc_sustained"],
                "mma_flops_per_algorithmic_flop": 4 if args.precision == "tc_split" else 1,
                "traffic": (None if (nt is None or nt.get("stale")) else nt.get("dram_bytes_per_launch")),
                "traffic_detail": nt,
                "conv_ms_per_step": tot_ms / psteps, "conv_launches_per_step": tot_n // psteps,
                "conv_share_of_step": (tot_ms / psteps) / ms_step, "per_class": per_class,
                "note": "event-bracketed launches run isolated (no programmatic overlap with their neighbours): "
                        "achieved / frac are lower bounds of what the same kernels do inside the CUDA-graph step",
                "step_frac_of_conv_flop_roofline": (value / world) * O.train_flops_per_image(cfg["H"], cfg["W"]) / (peak * 1e12)}

    # ---------------- side by side: (img/s, parity) of the two tensor-core modes, N=1 only ----------------
    # bf16 operands: the throughput mode (pose ~5e-2 off the fp32 reference: the format, not the kernels).
    # tc_split: the same tcgen05 engines on fp16 hi/lo operand planes, 4 MMAs per product -- meets the north-star
    # 1e-4 bar (tests/test_gpu_step.py::test_step_tc_split_strict*).  Both are measured here, in this run.
    modes = None
    if rank == 0 and world == 1 and not args.no_modes and args.precision in ("bf16", "tc_split"):
        other = "tc_split" if args.precision == "bf16" else "bf16"
        torch.cuda.synchronize()
        net2 = PoseNet(torchvision.models.resnet34(weights=None), droprate=args.droprate, pretrained=False,
                       filter_nans=(cfg["kind"] == "online"), precision=other, seed=7)
        model2 = net2 if cfg["kind"] == "posenet" else MapNet(net2)
        crit2 = make_crit()
        opt2 = Optimizer(params=[{"params": model2.parameters()}, {"params": list(crit2.parameters())}], method="adam",
                         base_lr=cfg["lr"], weight_decay=cfg["wd"])
        model2.cuda(); crit2.cuda(); model2.train()
        from geomapnet_b200.graph import GraphedTrainStep
        g2 = GraphedTrainStep(model2, crit2, opt2, dev_batches[0][0], dev_batches[0][1], max_grad_norm=cfg["clip"], warmup=2)
        for i in range(3):
            g2(*dev_batches[i % nb])
        torch.cuda.synchronize()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nsteps = max(5, args.steps // 2)
        m0.record()
        for i in range(nsteps):
            g2(*dev_batches[i % nb])
        m1.record()
        torch.cuda.synchronize()
        oms = m0.elapsed_time(m1) / nsteps
        parity = {"bf16": "loss <= 7e-3, pose <= 7e-2 relative to the fp32 reference at the BASELINE sizes (bf16 operands and "
                          "storage; tests/test_gpu_step.py::test_step_bf16_tensor_core)",
                  "tc_split": "loss <= 1.2e-5, pose <= 4.4e-5 relative to the fp32 reference on all 10 step goldens incl. the "
                              "BASELINE sizes: inside the north-star 1e-4 (tests/test_gpu_step.py::test_step_tc_split_strict*)"}
        modes = {args.precision: {"images_per_s": value, "ms_per_step": ms_step, "parity_vs_reference": parity[args.precision]},
                 other: {"images_per_s": frames(cfg) / (oms / 1000.0), "ms_per_step": oms, "steps": nsteps,
                         "parity_vs_reference": parity[other]},
                 "note": "same workload, same run, CUDA-graph step, inputs resident in HBM; fp32 CUDA-core engine "
                         "(precision fp32, also 1e-4): 957 img/s (profiles/r02c_bench_posenet_bs64_fp32.json)"}
        del g2, model2, net2

    # ---------------- cpu baseline: oracle port on the host cores (rank 0, N=1 only) -------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import weights
        cores = tune_cpu_threads(cfg["kind"])
        st = weights.make_state(7)
        sv = dict(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0)
        ref_frames = args.ref_frames if args.ref_frames else frames(cfg)
        n_tuples = min(cfg["N"], max(1, ref_frames // cfg["T"]))
        scfg = dict(cfg, N=n_tuples)
        cb = make_batches(scfg, 6, 300)
        ts = []
        tr = O.OracleTrainer(cfg["kind"], st, sv, lr=cfg["lr"], weight_decay=cfg["wd"], max_grad_norm=cfg["clip"],
                             droprate=args.droprate)
        for i, (x, targ) in enumerate(cb):
            t0 = time.time()
            tr.step(x, targ)
            if i >= 1:
                ts.append(time.time() - t0)
        cpu = {"value": frames(scfg) / (sum(ts) / len(ts)), "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "%d timed %d-frame steps of %s after 1 warm-up (oracle port, torch CPU ops)"
                         % (len(ts), frames(scfg), args.workload)}

    if rank == 0:
        line = {
            "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": {"bf16": "bf16", "bf16_simt": "bf16", "tc_split": "f16x2"}.get(args.precision, "f32"),
            "data": "synthetic",
            "config": {"workload": args.workload, "model": "PoseNet/MapNet ResNet-34", "frames_per_gpu": frames(cfg),
                       "global_frames": world * frames(cfg), "image": "%dx%d" % (cfg["H"], cfg["W"]),
                       "criterion": cfg["kind"], "optimizer": "adam (fused, flat)", "droprate": args.droprate,
                       "parallelism": "dp%d" % world, "precision": args.precision,
                       "cuda_graph": bool(args.graph), "eager_ms_per_step": eager_ms,
                       "l2": "per-step working set (activations + gradients, >3 GB) exceeds the 126 MB L2; "
                             "%d distinct input batches rotate" % nb},
            "e2e": {"value": e2e_val, "unit": "images/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "precision_modes": modes,
        }
        if world > 1 and args.workload != "posenet_bs64":
            # the default workload differs between N = 1 (posenet_bs64, BASELINE configs[1]) and N > 1 (configs[3]): the
            # weak-scaling baseline of THIS workload is its own one-GPU line, not the N = 1 default line
            n1 = {"mapnet_n32t3": (19172.4, "profiles/r02c_bench_mapnet_n32t3.json"),
                  "mapnetpp_n16t10": (20499.1, "profiles/r02c_bench_mapnetpp_n16t10.json")}.get(args.workload)
            if n1 is not None and args.precision == "bf16":
                line["config"]["one_gpu_same_workload"] = {
                    "images_per_s": n1[0], "source": n1[1],
                    "note": "bench.py --gpus 1 defaults to posenet_bs64; weak-scaling efficiency of this line = "
                            "value / (n_gpus x images_per_s), with `python bench.py --workload %s` as the N = 1 run"
                            % args.workload}
        print(json.dumps(line), flush=True)
    if world > 1:
        # the measurement is complete and printed: a communicator teardown that stalls (CUDA graphs holding captured NCCL
        # work) must not turn a finished run into a timeout
        def _bail():
            time.sleep(45)
            os._exit(0)
        threading.Thread(target=_bail, daemon=True).start()
        # CUDA graphs that captured NCCL work must be gone before their communicator is torn down
        step = gstep = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: posenet_bs64 on one GPU (BASELINE configs[1]), mapnet_n32t3 per GPU for --gpus > 1 "
                         "(BASELINE configs[3])")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16_simt", "tc_split"])
    ap.add_argument("--droprate", type=float, default=0.5)       # every reference .ini uses 0.5
    ap.add_argument("--ref-frames", type=int, default=0,
                    help="frames per CPU step; 0 = the full workload (a 64-frame step takes ~1.3 s on 16 cores)")
    ap.add_argument("--no-modes", action="store_true", help="skip the side-by-side strict-mode (tc_split) measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="run the step eagerly instead of replaying it as CUDA graphs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        print("warning: --gpus %d but WORLD_SIZE=%d; using WORLD_SIZE" % (args.gpus, world), file=sys.stderr)
    if args.gpus > 1 and world == 1 and args.impl == "b200":
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    if args.workload is None:
        args.workload = "posenet_bs64" if max(world, args.gpus) == 1 else "mapnet_n32t3"
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args, cfg, rank, world)
        return
    args.steps = args.steps if args.steps is not None else 20
    args.warmup = args.warmup if args.warmup is not None else 5
    if args.warmup < 3:
        args.warmup = 3
    run_b200(args, cfg, rank, local_rank, world)


if __name__ == "__main__":
    main()
