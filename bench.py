#!/usr/bin/env python
"""bench.py — BASELINE.json metric: 3-D vol-pairs/sec (160x192x224) for one diffeomorphic VxmDense TRAINING step
(forward + NCC/Grad losses + backward + gradient allreduce + Adam) at N GPUs, one volume pair per GPU per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL = (160, 192, 224)
METRIC = "vol-pairs/sec (3D 160x192x224 VxmDense int_steps=7 train step, NCC+Grad, Adam)"
UNIT = "vol-pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", type=int, nargs=3, default=list(FULL), help="debug only; the metric is quoted at 160 192 224")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine", default=None, choices=["f32", "bf16"], help="convolution engine (default: bf16 tensor-core engine)")
    ap.add_argument("--no-graph", action="store_true", help="time eager kernel launches instead of the captured CUDA graph")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline legs")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# conv FLOPs of the default U-Net per training step
# ------------------------------------------------------------------------------------------------
def conv_flops_per_step(shape):
    import numpy as np
    V = float(np.prod(shape))
    layers = [  # (cin, cout, voxel fraction, has_dgrad)
        (2, 16, 1, False), (16, 32, 1 / 8, True), (32, 32, 1 / 64, True), (32, 32, 1 / 512, True),
        (32, 32, 1 / 4096, True), (64, 32, 1 / 512, True), (64, 32, 1 / 64, True), (64, 32, 1 / 8, True),
        (48, 32, 1, True), (32, 16, 1, True), (16, 16, 1, True), (16, 3, 1, True)]
    fwd = sum(2 * 27 * ci * co * V * f for ci, co, f, _ in layers)
    bwd = sum(2 * 27 * ci * co * V * f * (2 if dg else 1) for ci, co, f, dg in layers)
    return fwd, fwd + bwd


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken in [t0, t1] (host clock).  The sampler is started before the warm-up so that it is
        already streaming; when the timed region is shorter than the sampling period the nearest sample taken under the
        same load (warm-up steps run back to back with the timed ones) is used and `window` says so."""
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.06)   # let the sample that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        rows, window = list(self.rows), "timed region"
        if t0 is not None:
            inside = [r for r in rows if t0 <= r[0] <= t1 + 0.06]
            if inside:
                rows = inside
            elif rows:
                rows = [min(rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))]
                window = "nearest sample under load (timed region shorter than the 50 ms sampling period)"
        for _, r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                    reasons=sorted(reasons), window=window)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's torch CPU path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_time(shape, steps, warmup, budget_s):
    """Time the oracle restatement of the reference training step (oracle/ref_torch.py: same torch CPU operators
    the reference calls) on all host cores.  Returns (sec per FULL-SIZE pair, cores, sample description)."""
    import numpy as np
    import torch
    from oracle import cases, ref_torch
    cores = os.cpu_count() or 1

    def make(shp):
        cfg = dict(inshape=tuple(shp), nb_unet_features=None, nb_unet_levels=None, unet_feat_mult=1, nb_unet_conv_per_level=1,
                   int_steps=7, int_downsize=2, bidir=False, use_probs=False, src_feats=1, trg_feats=1, unet_half_res=False)
        sd = {k: v.requires_grad_(True) for k, v in ref_torch.init_state_dict(cfg, seed=1234, flow_std=1e-2).items()}
        opt = torch.optim.Adam(list(sd.values()), lr=1e-4)
        g = torch.Generator().manual_seed(1234)
        s = torch.rand((1, 1) + tuple(shp), generator=g)
        t = torch.rand((1, 1) + tuple(shp), generator=g)
        return cfg, sd, opt, s, t

    def run(shp, n):
        cfg, sd, opt, s, t = make(shp)
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            ref_torch.train_step(sd, cfg, opt, s, t, image_loss="ncc", lam=0.01)
            ts.append(time.perf_counter() - t0)
        return ts

    sub = tuple(max(16, (d // 2 // 16) * 16) for d in shape)
    frac = float(np.prod(sub)) / float(np.prod(shape))
    # give the CPU path its best thread count (oneDNN / ATen do not always scale to every core of a big host)
    best = None
    for nt in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16)}, reverse=True):
        torch.set_num_threads(nt)
        tt = run(sub, 2)[-1]
        if best is None or tt < best[0]:
            best = (tt, nt)
    t_sub, cores_used = best
    torch.set_num_threads(cores_used)
    cores = cores_used
    est_full = t_sub / frac
    if (steps + warmup) * est_full <= budget_s:
        ts = run(shape, steps + warmup)[warmup:]
        return sum(ts) / len(ts), cores, "%d full-size %s steps after %d warm-up, torch %s CPU fp32, %d threads" % (
            len(ts), "x".join(map(str, shape)), warmup, torch.__version__, cores)
    n = max(1, min(steps, int(budget_s / max(t_sub, 1e-3)) - warmup))
    ts = run(sub, n + max(1, min(warmup, 2)))[max(1, min(warmup, 2)):]
    per_full = (sum(ts) / len(ts)) / frac
    return per_full, cores, ("%d steps on a %s sub-volume (%.3f of the voxels; time scaled by 1/%.3f), torch %s CPU fp32, "
                             "%d threads" % (len(ts), "x".join(map(str, sub)), frac, frac, torch.__version__, cores))


def reference_arm(args):
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, cores, sample = cpu_step_time(tuple(args.shape), args.steps, args.warmup, budget_s=150.0)
    v = 1.0 / sec
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload="3D %s VxmDense diffeomorphic (int_steps=7, int_downsize=2), NCC+0.01*Grad, Adam, batch 1"
                            % "x".join(map(str, args.shape)), note="reference torch CPU path (oracle port) on host cores"),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def b200_arm(args):
    import numpy as np
    import torch
    import voxelmorph_b200 as vxm
    from voxelmorph_b200 import dist as vdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)")
    world, rank, local = vdist.init_from_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus %d needs torchrun (one process per GPU)" % args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    vxm._lib.load()
    os.environ["VXM_B200_CONV_ENGINE"] = args.engine or os.environ.get("VXM_B200_CONV_ENGINE", "bf16")
    peaks = load_peaks()
    shape = tuple(args.shape)
    V = int(np.prod(shape))

    # ---- model, optimizer, data ----------------------------------------------------------------
    torch.manual_seed(1234)
    model = vxm.networks.VxmDense(inshape=shape, int_steps=7, int_downsize=2)
    with torch.no_grad():
        model.flow.weight.normal_(0, 1e-2)   # trained-like flow scale so that warps / VecInt do real work
    model.to(dev).train()
    opt = vxm.optim.FusedAdam(model.parameters(), lr=1e-4, world_size=world)
    vdist.broadcast_params(opt.fp.flat)
    ncc = vxm.losses.NCC().loss
    grad = vxm.losses.Grad("l2", loss_mult=2).loss

    # synthetic pairs (generated on the device by torch, seeds differ per rank): smooth volume + warped copy
    NPAIR = 4
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    pairs_host = []
    st = vxm.layers.SpatialTransformer(shape)
    for i in range(NPAIR):
        coarse = torch.rand((1, 1, shape[0] // 8, shape[1] // 8, shape[2] // 8), generator=g, device=dev)
        src = torch.nn.functional.interpolate(coarse, size=shape, mode="trilinear", align_corners=True)
        src = (src + 0.05 * torch.rand(src.shape, generator=g, device=dev)).clamp_(0, 1).contiguous()
        fl = torch.nn.functional.interpolate(torch.randn((1, 3, shape[0] // 16, shape[1] // 16, shape[2] // 16), generator=g,
                                                         device=dev) * 3.0, size=shape, mode="trilinear", align_corners=True)
        trg = st(src, fl.contiguous())
        pairs_host.append((src.cpu().pin_memory(), trg.cpu().pin_memory()))
    pairs_dev = [(s.to(dev), t.to(dev)) for s, t in pairs_host]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def eager_step(S, T):
        opt.zero_grad()
        y, flow = model(S, T)
        loss = ncc(T, y) + 0.01 * grad(None, flow)
        loss.backward()
        vdist.allreduce_grads(opt.fp.grad)
        opt.step()
        return loss

    # the whole step (zero-grad, fwd, losses, bwd, allreduce, Adam) captured once in a CUDA graph and replayed
    step, graphed = eager_step, False
    launches_per_step = None
    if not args.no_graph:
        from voxelmorph_b200.trainer import GraphedTrainStep
        try:
            n0 = vxm._lib.launch_count()
            eager_step(*pairs_dev[0])
            launches_per_step = vxm._lib.launch_count() - n0
            trainer = GraphedTrainStep(model, opt, image_loss="ncc", lam=0.01, int_downsize=2).capture(*pairs_dev[0])
            step, graphed = trainer, True
        except Exception as e:  # noqa: BLE001 - report and fall back to eager launches
            print("bench.py: CUDA graph capture failed (%s); timing eager launches" % e, file=sys.stderr)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------
    W, K = max(3, args.warmup), args.steps
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    for i in range(W):
        step(*pairs_dev[i % NPAIR])
    barrier()

    # ---- device-resident timed region ------------------------------------------------------------
    n0 = vxm._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.time()
    e0.record()
    for i in range(K):
        step(*pairs_dev[i % NPAIR])
    e1.record()
    barrier()
    t_host1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = (launches_per_step * K) if graphed else (vxm._lib.launch_count() - n0)
    ms = vdist.max_over_ranks(ms, dev)
    clk = clocks.stop(t_host0, t_host1) if rank == 0 else None
    value = world * K / (ms * 1e-3)

    # ---- end-to-end: host buffers, H2D of the pair + D2H of the loss inside the timed region --------
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(pairs_dev[0][0]), torch.empty_like(pairs_dev[0][1])) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])
            bufs[b][0].copy_(pairs_host[i % NPAIR][0], non_blocking=True)
            bufs[b][1].copy_(pairs_host[i % NPAIR][1], non_blocking=True)
            ready[b].record(copy_stream)

    for b in range(2):
        freed[b].record()
    barrier()
    e0.record()
    prefetch(0)
    for i in range(K):
        if i + 1 < K:
            prefetch(i + 1)
        b = i % 2
        torch.cuda.current_stream().wait_event(ready[b])
        loss = step(*bufs[b])
        freed[b].record()
        loss_host.copy_(loss.detach(), non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = vdist.max_over_ranks(e0.elapsed_time(e1), dev)
    e2e = dict(value=world * K / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=2 * V * 4, d2h_bytes_per_step=4,
               ms_per_step=ms_e2e / K, api="voxelmorph_b200.networks.VxmDense + losses.NCC/Grad + optim.FusedAdam, pinned host "
               "buffers, H2D double-buffered on a copy stream")

    # ---- roofline of the dominant kernel family (Conv3d) measured live with CUDA events -------------
    # One eager step records every convolution launch (function + arguments); the recorded launches are then
    # re-issued back to back between two CUDA events on the launching stream, behind a device-side sleep so that the
    # host runs ahead and the events bracket kernel time only (no launch gaps).
    from voxelmorph_b200 import ops
    from voxelmorph_b200 import tc as tcmod
    calls = []

    def recording(fn):
        def inner(*a, **k):
            calls.append((fn, a, k))
            return fn(*a, **k)
        return inner

    o_cf, o_cw, o_ct = tcmod.conv_fwd, tcmod.conv_wgrad, tcmod.conv_fwd_t
    orig_fwd, orig_bwd = ops._ConvK3Fn.forward, ops._ConvK3Fn.backward
    conv_total_ms = None
    if ops.conv_engine() == "bf16":
        tcmod.conv_fwd, tcmod.conv_wgrad, tcmod.conv_fwd_t = recording(o_cf), recording(o_cw), recording(o_ct)
        eager_step(*pairs_dev[0])
        tcmod.conv_fwd, tcmod.conv_wgrad, tcmod.conv_fwd_t = o_cf, o_cw, o_ct
        torch.cuda.synchronize()
        reps = []
        for _ in range(3):
            torch.cuda._sleep(12000000)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for fn, a, k in calls:
                fn(*a, **k)
            c1.record()
            torch.cuda.synchronize()
            reps.append(c0.elapsed_time(c1))
        conv_total_ms = statistics.median(reps)
        n_conv_launches = len(calls)
        if os.environ.get("VXM_BENCH_VERBOSE") and rank == 0:
            for fn, a, k in calls:      # per-launch breakdown (stderr)
                torch.cuda._sleep(2000000)
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(); fn(*a, **k); c1.record(); torch.cuda.synchronize()
                shp = [tuple(x.shape) for x in a[:3] if hasattr(x, "shape")]
                print("  %-12s %8.1f us  %s" % (fn.__name__, c0.elapsed_time(c1) * 1e3, shp), file=sys.stderr)
        del calls
    else:
        conv_ms = []

        def timed(fn):
            def inner(*a, **k):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                r = fn(*a, **k)
                a1.record()
                conv_ms.append((a0, a1))
                return r
            return inner

        ops._ConvK3Fn.forward = staticmethod(timed(orig_fwd))
        ops._ConvK3Fn.backward = staticmethod(timed(orig_bwd))
        eager_step(*pairs_dev[0])
        torch.cuda.synchronize()
        ops._ConvK3Fn.forward, ops._ConvK3Fn.backward = staticmethod(orig_fwd), staticmethod(orig_bwd)
        conv_total_ms = sum(x.elapsed_time(y) for x, y in conv_ms)
        n_conv_launches = len(conv_ms)
    if rank != 0:
        _leave(world, rank)
        return
    _, flops_step = conv_flops_per_step(shape)
    ach = flops_step / (conv_total_ms * 1e-3) / 1e12
    engine = ops.conv_engine()
    # DRAM traffic of the conv family per step: taken from the committed ncu pass over this same command
    # (profiles/r1_traffic.json; ncu cannot run inside a timed bench), valid for the full-size bf16 workload only.
    traffic, traffic_src = None, None
    tj = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")
    if engine == "bf16" and tuple(shape) == (160, 192, 224) and os.path.exists(tj):
        with open(tj) as f:
            tinfo = json.load(f)
        traffic, traffic_src = tinfo["conv_dram_mbytes_per_step"] * 1e6, tinfo["source"]
    roofline = dict(bound="tensor", kernel="conv3d k3 fwd+dgrad+wgrad, all 12 layers (%s)" % ("tcgen05 bf16 implicit GEMM" if engine == "bf16" else "fp32 FFMA engine"),
                    achieved=ach, peak=peaks["tf_sus"], unit="TFLOP/s", frac=ach / peaks["tf_sus"], traffic=traffic, traffic_unit="bytes per step (all conv launches)", traffic_source=traffic_src,
                    peak_source=peaks["source"] + ", sustained bf16", ms_per_step=conv_total_ms, conv_launches=n_conv_launches,
                    share_of_step=conv_total_ms / (ms / K), flops_per_step=flops_step)

    kernels = {} if args.no_kernels else kernel_rooflines(vxm, dev, shape, peaks)

    # ---- CPU baseline (oracle port of the reference's torch CPU path) -------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        sec, cores, sample = cpu_step_time(shape, 1, 0, budget_s=30.0)
        cpu = dict(value=1.0 / sec, unit=UNIT, cores=cores, kind="port", sample=sample)

    act_gb = 4.0 * V * (2 + 16 + 48 + 32 + 16 + 16 + 3) / 1e9
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32" if engine == "f32" else "bf16 (conv operands) / f32 (accumulate, warp, VecInt, losses)",
                data="synthetic", impl="b200",
                config=dict(workload="3D %s VxmDense diffeomorphic (int_steps=7, int_downsize=2), default U-Net features, "
                            "NCC(9^3)+0.01*Grad(l2), Adam lr 1e-4, 1 pair per GPU" % "x".join(map(str, shape)),
                            global_batch=world, parallelism="dp%d (one flat-gradient allreduce per step)" % world,
                            conv_engine=engine, cuda_graph=graphed,
                            l2="inputs rotate over %d resident pairs; per-step working set ~%.1f GB of full-resolution "
                               "activations >> 126 MB L2, so no explicit flush" % (NPAIR, act_gb)),
                clocks=clk, e2e=e2e, gpu_launches=int(launches), launches_per_step=launches / K,
                roofline=roofline, kernels=kernels, cpu_baseline=cpu)
    print(json.dumps(line), flush=True)
    _leave(world, rank)


def _leave(world, rank=None):
    """End of a rank's work.  Under torchrun every rank leaves with os._exit(0), rank 0 last: the captured CUDA graph
    still holds NCCL kernels, and tearing the process group down with it alive (destroy_process_group / interpreter
    shutdown) blocked both ranks after the JSON line had been printed (2 x B200, round 1).  The ranks meet on the
    rendezvous store (no collective): rank 0 posts `bench_done`, the others acknowledge, then everybody exits."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world <= 1:
        return
    import datetime
    import torch
    import torch.distributed as dist
    try:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        store = dist.distributed_c10d._get_default_store()
        store.set_timeout(datetime.timedelta(seconds=1800))
        if rank is None:
            rank = dist.get_rank()
        if rank == 0:
            store.set("bench_done", "1")
            t0 = time.time()
            while store.add("bench_ack", 0) < world - 1 and time.time() - t0 < 60.0:
                time.sleep(0.05)
        else:
            store.wait(["bench_done"])      # rank 0 may still be timing its CPU baseline
            store.add("bench_ack", 1)
            time.sleep(0.2)
    except Exception as e:  # noqa: BLE001 - leaving must not fail
        print("bench.py: exit rendezvous skipped (%s)" % e, file=sys.stderr)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def kernel_rooflines(vxm, dev, shape, peaks):
    """Each memory-bound kernel timed alone (CUDA events, 3 warm-up + 10 timed launches, 256 MB L2 flush between
    launches); algorithmic bytes per SURVEY.md section 8(d)."""
    import numpy as np
    import torch
    V = int(np.prod(shape))
    half = tuple(s // 2 for s in shape)
    Vh = int(np.prod(half))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    out = {}

    def timeit(fn, nbytes, name, note=""):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            torch.cuda._sleep(400000)   # keep the GPU busy while the host enqueues, so the events bracket the kernel only
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = statistics.median(ts)
        gbs = nbytes / (t * 1e-3) / 1e9
        out[name] = dict(bound="hbm", achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"], us=t * 1e3,
                         algorithmic_bytes=nbytes, note=note)

    src = torch.rand((1, 1) + shape, device=dev)
    smooth = lambda shp, sig: (torch.nn.functional.interpolate(  # noqa: E731  registration-like smooth displacement field
        torch.randn((1, 3) + tuple(max(2, s // 16) for s in shp), device=dev) * sig, size=shp, mode="trilinear",
        align_corners=True).contiguous())
    flow = smooth(shape, 3.0)
    rough = torch.randn((1, 3) + shape, device=dev) * 3.0
    st = vxm.layers.SpatialTransformer(shape)
    stn = vxm.layers.SpatialTransformer(shape, mode="nearest")
    vel = smooth(half, 2.0)
    vi = vxm.layers.VecInt(half, 7)
    down, up = vxm.layers.ResizeTransform(2, 3), vxm.layers.ResizeTransform(0.5, 3)
    I, J = torch.rand((1, 1) + shape, device=dev), torch.rand((1, 1) + shape, device=dev)
    ncc = vxm.losses.NCC().loss
    with torch.no_grad():
        timeit(lambda: st(src, flow), V * 20, "warp_fwd_linear", "C=1, smooth flow sigma=3 voxels (registration-like)")
        timeit(lambda: stn(src, flow), V * 20, "warp_fwd_nearest", "C=1, smooth flow; exact replay of the reference's coordinate arithmetic (bit-exact labels)")
        timeit(lambda: st(src, rough), V * 20, "warp_fwd_linear_white_noise_flow", "C=1, i.i.d. N(0,3^2) flow per voxel (worst-case gather locality)")
        timeit(lambda: vi(vel), Vh * 24 * 7, "vecint_fwd_7steps", "single cooperative launch; field is L2 resident, so frac can "
               "exceed 1 against the HBM peak")
        timeit(lambda: down(flow), (V + Vh) * 12, "resize_down")
        timeit(lambda: up(vel), (V + Vh) * 12, "resize_up")
        timeit(lambda: ncc(I, J), V * 8, "ncc_fwd", "no saved fields (inference / validation)")

    # backward legs: the autograd node's backward is timed alone (forward outside the events); bytes per SURVEY 8(d)
    def timeit_bwd(make, nbytes, name, note=""):
        def run():
            y, g = make()
            return lambda: y.backward(g, retain_graph=True)
        fn = run()
        timeit(fn, nbytes, name, note)

    fl = flow.clone().requires_grad_(True)
    timeit_bwd(lambda: (st(src, fl), torch.ones((1, 1) + shape, device=dev)), V * 36, "warp_bwd_linear",
               "d/d flow only (the moving image needs no gradient in training): reads grad, flow, src; writes 3 planes")
    vl = vel.clone().requires_grad_(True)
    timeit_bwd(lambda: (vi(vl), torch.ones_like(vel)), Vh * 36 * 7, "vecint_bwd_7steps", "single cooperative launch, red.global.add.v4 scatter")
    timeit_bwd(lambda: (up(vl), torch.ones((1, 3) + shape, device=dev)), (V + Vh) * 12, "resize_up_bwd", "adjoint of the x2 upsampling (reads the full-resolution gradient)")
    timeit_bwd(lambda: (down(fl), torch.ones((1, 3) + half, device=dev)), (V + Vh) * 12, "resize_down_bwd")
    Jg = J.clone().requires_grad_(True)
    timeit(lambda: ncc(I, Jg), V * 8 + V * 12, "ncc_fwd_training", "forward that also stores the 3 fields the backward box-filters")
    timeit_bwd(lambda: (ncc(I, Jg), torch.ones((), device=dev)), V * 12 + V * 12, "ncc_bwd", "reads I, J + 3 saved fields, writes dJ")
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        b200_arm(a)
