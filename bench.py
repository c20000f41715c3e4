#!/usr/bin/env python
"""Benchmark of the SPNet hot path on B200: Xception-SPNet full training step
(forward + YOLO-ellipse loss + backward + Keras Adam), bf16, batch 64 per GPU, 384x512x1
gen_fake_espi-style synthetic frames (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0). N > 1 is launched by torchrun (one rank per GPU, NCCL).
`value`   : images/s with the batch already resident in HBM (CUDA-graph replay, device timed).
`e2e`     : images/s through the engine's public input path with HOST buffers: every step's inputs are
            copied from pinned host memory inside the timed region (prefetched one step ahead on a copy
            stream, exactly as SPNetModel.fit does) and the loss is read back to the host every step.
`roofline`: the dominant kernel family of the step (its launches of one step replayed as a CUDA graph
            and timed with CUDA events); `roofline_other` holds the other families.
`--impl reference`: the reference network's CPU path. TensorFlow 1.14 / Keras 2.1.3 cannot run in
this image, so this arm times the oracle restatement (kind "port") of the same
training step on the box's host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, N_OUT = 384, 512, 576
BATCH_PER_GPU = 64
METRIC = "train images/sec (Xception-SPNet, 512x384)"
LR = 4e-5


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_pool(n, seed0):
    from spnet_b200 import fake_espi
    X, Y, _ = fake_espi.make_dataset(n, base_seed=seed0)
    return X, Y


# -------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """CPU arm: oracle restatement of the same training step on the host cores (kind 'port')."""
    if rank != 0:
        return
    import torch
    from oracle import xception_torch as xt
    ncores = os.cpu_count() or 1
    torch.set_num_threads(ncores)
    Bs = 4  # bounded sample of the batch-64 workload: 4 images per step
    X, Y = make_pool(Bs, 10_000)
    spec = xt.xception_spnet_spec(H, W, N_OUT)
    model = xt.OracleSPNet(xt.init_weights(spec, seed=1), H, W)
    times = []
    for i in range(args.warmup_ref + args.steps_ref):
        t0 = time.perf_counter()
        _, _, _, grads = model.loss_and_grads(X, Y)
        model.adam_step(grads, LR)
        dt = time.perf_counter() - t0
        if i >= args.warmup_ref:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = Bs / (ms / 1e3)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
           "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "Xception-SPNet train step fwd+bwd+loss+Adam, 384x512x1, batch 64/GPU",
                      "note": "reference CPU path = oracle restatement (TF1.14/Keras2.1.3 not runnable here)"},
           "cpu_baseline": {"value": val, "unit": "images/s", "cores": ncores, "kind": "port",
                            "sample": "%d steps of %d images (of the batch-64 workload), torch %s CPU fp32" % (
                                args.steps_ref, Bs, torch.__version__)},
           "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# -------------------------------------------------------------------------------------------------
def kernel_breakdown(eng, lib, steps=2):
    """Eager steps with a CUDA-event pair around every kernel launch -> per-family device time."""
    import torch
    lib.profile = []
    for _ in range(steps):
        eng.train_step(LR)
    torch.cuda.synchronize()
    prof, lib.profile = lib.profile, None
    fam = {}
    for name, args, e0, e1 in prof:
        d = fam.setdefault(name, {"ms": 0.0, "launches": 0, "items": []})
        ms = e0.elapsed_time(e1)
        d["ms"] += ms
        d["launches"] += 1
        d["items"].append((args, ms))
    return fam, steps


def algorithmic_work(name, a):
    """(bytes, flops) one launch must move / compute, from its C-ABI arguments."""
    def es(dtype):
        return 2 if dtype == 1 else 4
    a = tuple(a)
    if name == "spnet_dwconv3x3_fwd":      # in,k,a,b,relu,out,dtype,B,H,W,C,stream
        n = a[7] * a[8] * a[9] * a[10]
        return 2 * n * es(a[6]), 18 * n
    if name == "spnet_dwconv3x3_bwd_fused":  # gout,in,k,a,b,relu,mean,rstd,stats,add,adds,gin,dk,dtype,B,H,W,C
        n = a[14] * a[15] * a[16] * a[17]
        t = 3 + (1 if a[9] else 0)
        return t * n * es(a[13]), 36 * n
    if name == "spnet_gemm_bf16":          # A,lda,amn,B,ldb,bmn,D,ldd,mode,M,N,K,splits,...
        M, N, K = a[9], a[10], a[11]
        return 2 * (M * K + N * K) + (2 if a[8] == 0 else 4) * M * N, 2 * M * N * K
    return 0, 0


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    from spnet_b200._lib import lib as get_lib
    from spnet_b200 import engine as engine_mod
    from spnet_b200 import multi_gpu
    # the headline workload is Xception (BASELINE configs[1]); --backbone runs the same step on the other
    # backbones of BASELINE configs[2] / [3] for the record (profiles/README.md), not as the bench line
    Engine = {"Xception": engine_mod.XceptionSPNetEngine, "MobileNet": engine_mod.MobileNetSPNetEngine,
              "InceptionResNetV2": engine_mod.InceptionResNetV2SPNetEngine}[args.backbone]
    global METRIC
    METRIC = METRIC.replace("Xception", args.backbone)

    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = get_lib()
    B = BATCH_PER_GPU
    npool = 2 * B
    X, Y = make_pool(npool, 1_000_000 * rank)
    Xp = torch.from_numpy(X).pin_memory()
    Yp = torch.from_numpy(Y).pin_memory()
    Xd, Yd = Xp.to(dev), Yp.to(dev)

    eng = Engine(H, W, B, dtype="bf16", device=str(dev), seed=1)
    if world > 1:
        multi_gpu.attach_data_parallel(eng)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # eager warm-up (also counts kernel launches per step), then capture the step into a CUDA graph
    eng.x0.copy_(Xd[:B]); eng.y_true.copy_(Yd[:B])
    l0 = lib.launches
    eng.train_step(LR)
    torch.cuda.synchronize()
    launches_per_step = lib.launches - l0
    eng.capture()
    for i in range(args.warmup):
        o = (i % 2) * B
        eng.x0.copy_(Xd[o:o + B]); eng.y_true.copy_(Yd[o:o + B])
        eng.train_step(LR)
    barrier()

    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        o = (i % 2) * B
        eng.x0.copy_(Xd[o:o + B]); eng.y_true.copy_(Yd[o:o + B])
        eng.train_step(LR)
    ev1.record()
    barrier()
    ms_dev = ev0.elapsed_time(ev1)

    # the box's pinned host->device rate for one batch of frames, measured alone (not part of any timed region):
    # the end-to-end figure below cannot exceed batch / (this copy's time) and single-GPU boxes differ a lot here
    hb0, hb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch_dev = torch.empty_like(Xd[:B])
    scratch_dev.copy_(Xp[:B], non_blocking=True)
    torch.cuda.synchronize()
    hb0.record()
    for _ in range(3):
        scratch_dev.copy_(Xp[:B], non_blocking=True)
    hb1.record()
    torch.cuda.synchronize()
    h2d_gbps = 3 * scratch_dev.numel() * 4 / (hb0.elapsed_time(hb1) * 1e-3) / 1e9
    del scratch_dev

    # ---- timed region 2: end to end from host buffers (pinned H2D + loss D2H every step)
    loss_host = torch.zeros(6).pin_memory()
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    last = None
    eng.prefetch_batch(Xp[:B], Yp[:B])  # the first batch's H2D copy is inside the timed region too
    for i in range(args.steps):
        eng.take_prefetched()
        if i + 1 < args.steps:
            o = ((i + 1) % 2) * B
            eng.prefetch_batch(Xp[o:o + B], Yp[o:o + B])  # next batch travels while this step computes (as Model.fit does)
        loss6 = eng.train_step(LR)
        loss_host.copy_(loss6, non_blocking=False)
        last = float(loss_host[0])
    ev3.record()
    barrier()
    ms_e2e = ev2.elapsed_time(ev3)
    clocks = sampler.stop()

    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return

    total_imgs = B * world * args.steps
    value = total_imgs / (ms_dev / 1e3)
    e2e = total_imgs / (ms_e2e / 1e3)
    peaks = load_peaks()

    # ---- per-family view of the step: eager steps with a CUDA-event pair around every launch give the
    #      share of each kernel family; the roofline numbers come from replaying each family's launches
    #      (same arguments, same buffers, step order) as one CUDA graph timed with CUDA events, so that
    #      the ~4 us per-launch cost of eager event pairs does not deflate kernels that run 10-30 us
    eng.graph = None
    hook, eng.grad_hook = eng.grad_hook, None  # time this rank's kernels only
    fam, psteps = kernel_breakdown(eng, lib)
    eng.grad_hook = hook
    total_ms = sum(d["ms"] for d in fam.values())
    share = sorted(((d["ms"] / total_ms, n, d["launches"] // psteps) for n, d in fam.items()), reverse=True)
    breakdown = [{"kernel": n, "share": round(s, 4), "launches_per_step": l} for s, n, l in share[:14]]

    def graph_us(calls, reps=4):
        """Average device time per launch of `calls` [(name, args)] replayed as one CUDA graph."""
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st = torch.cuda.current_stream().cuda_stream
            for name, a in calls:
                rc = getattr(lib._dll, name)(*(tuple(a[:-1]) + (st,)))
                if rc != 0:
                    raise RuntimeError("%s failed in the family replay: %s" % (name, lib.last_error()))
        g.replay()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            g.replay()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) * 1e3 / (reps * len(calls))

    RIDGE = peaks["tf_sust"] * 1e12 / (peaks["hbm"] * 1e9)  # FLOP/B above which a GEMM is tensor-bound
    one_step = {n: d["items"][:len(d["items"]) // psteps] for n, d in fam.items()}
    fams = {"gemm_tensor": ("tensor", "gemm_tc_kernel (tcgen05), shapes above the ridge (%.0f FLOP/B)" % RIDGE, []),
            "gemm_hbm": ("hbm", "gemm_tc_kernel (tcgen05), shapes below the ridge", []),
            "dwconv_fwd": ("hbm", "dw3x3_fwd_packed_kernel (+BN/ReLU on load)", []),
            "dwconv_bwd": ("hbm", "dw3x3_bwd_packed_kernel (dgrad+wgrad+mask+BN sums)", [])}
    for a, _ in one_step.get("spnet_gemm_bf16", []):
        b_, f_ = algorithmic_work("spnet_gemm_bf16", a)
        fams["gemm_tensor" if f_ / b_ >= RIDGE else "gemm_hbm"][2].append(("spnet_gemm_bf16", a))
    for a, _ in one_step.get("spnet_dwconv3x3_fwd", []):
        fams["dwconv_fwd"][2].append(("spnet_dwconv3x3_fwd", a))
    for a, _ in one_step.get("spnet_dwconv3x3_bwd_fused", []):
        fams["dwconv_bwd"][2].append(("spnet_dwconv3x3_bwd_fused", a))
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1", "traffic.json")))
    except Exception:
        traffic = {}
    roofs = {}
    for key, (bound, label, calls) in fams.items():
        if not calls:
            continue
        us = graph_us(calls)
        nb = sum(algorithmic_work(n, a)[0] for n, a in calls)
        nf = sum(algorithmic_work(n, a)[1] for n, a in calls)
        r = {"bound": bound, "kernel": label, "launches": len(calls), "avg_launch_ms": us * 1e-3,
             "traffic": traffic.get(key, {}).get("bytes_per_launch"), "traffic_source": traffic.get(key, {}).get("source"),
             "timing": "CUDA-graph replay of the family's launches of one step, CUDA events"}
        t_s = us * 1e-6 * len(calls)
        if bound == "hbm":
            ach = nb / t_s / 1e9
            r.update({"achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                      "algorithmic_bytes_per_step": nb})
        else:
            ach = nf / t_s / 1e12
            r.update({"achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sust"],
                      "algorithmic_flops_per_step": nf})
        r["share_of_step"] = t_s * 1e3 / (ms_dev / args.steps)
        roofs[key] = r
    # per-shape view of the GEMMs (M,N,K,a_mn,b_mn): eager-event time, TFLOP/s
    shapes = {}
    for a, ms in fam.get("spnet_gemm_bf16", {"items": []})["items"]:
        key = (a[9], a[10], a[11], a[2], a[5])
        d = shapes.setdefault(key, [0, 0.0])
        d[0] += 1
        d[1] += ms
    gemm_shapes = [{"M": k[0], "N": k[1], "K": k[2], "a_mn": k[3], "b_mn": k[4], "launches_per_step": v[0] // psteps,
                    "us_each_eager": round(1e3 * v[1] / v[0], 1), "tflops": round(2.0 * k[0] * k[1] * k[2] / (v[1] / v[0] * 1e-3) / 1e12, 1)}
                   for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1])[:10]]
    dominant = max(roofs.values(), key=lambda r: r["share_of_step"])
    other = [r for r in roofs.values() if r is not dominant]

    # ---- inference throughput (forward only, moving-stat BN), same batch / shape, CUDA-graph replay
    infer = None
    try:
        del eng
        torch.cuda.empty_cache()
        ieng = Engine(H, W, B, dtype="bf16", device=str(dev), seed=1, training=False)
        ieng.x0.copy_(Xd[:B])
        ieng.forward(training=False)
        torch.cuda.synchronize()
        ig = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ig):
            ieng.forward(training=False)
        for _ in range(3):
            ig.replay()
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for i in range(args.steps):
            ieng.x0.copy_(Xd[(i % 2) * B:(i % 2) * B + B])
            ig.replay()
        i1.record()
        torch.cuda.synchronize()
        ims = i0.elapsed_time(i1) / args.steps
        infer = {"value": B / (ims / 1e3), "unit": "images/s (per GPU)", "ms_per_batch": ims, "batch": B}
    except Exception as e:  # the training numbers above stay valid
        infer = {"error": str(e)[:200]}

    # ---- CPU baseline on this box's host cores: bounded sample of the same workload
    cpu = cpu_baseline()

    out = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": "%s-SPNet train step fwd+bwd+YOLO-ellipse loss+Keras Adam, 384x512x1, batch %d/GPU" % (args.backbone, B),
                      "global_batch": B * world, "parallelism": "dp%d" % world, "dropout": 0.1, "l2": 1e-4,
                      "input_pool": "%d distinct gen_fake_espi-style frames per rank" % npool,
                      "l2_cache": "per-step working set (activations, several GB) >> 126 MB L2; no explicit flush",
                      "cuda_graph": True, "loss_last": last},
           "clocks": clocks,
           "e2e": {"value": e2e, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                   "h2d_bytes_per_step": int(B * (H * W * 4 + N_OUT * 4)), "d2h_bytes_per_step": 24,
                   "h2d_gb_per_s_of_this_box": round(h2d_gbps, 2),
                   "h2d_bound_images_per_s": round(B * world / (B * H * W * 4 / (h2d_gbps * 1e9)), 1)},
           "gpu_launches": int(launches_per_step * args.steps * 2),
           "roofline": dominant, "roofline_other": other, "kernel_breakdown": breakdown, "eager_step_ms": total_ms / psteps, "gemm_shapes": gemm_shapes,
           "peaks": peaks, "inference": infer, "cpu_baseline": cpu}
    print(json.dumps(out), flush=True)


def cpu_baseline():
    import torch
    from oracle import xception_torch as xt
    ncores = os.cpu_count() or 1
    torch.set_num_threads(ncores)
    Bs = 4
    X, Y = make_pool(Bs, 10_000)
    model = xt.OracleSPNet(xt.init_weights(xt.xception_spnet_spec(H, W, N_OUT), seed=1), H, W)
    ts = []
    for i in range(3):
        t0 = time.perf_counter()
        _, _, _, grads = model.loss_and_grads(X, Y)
        model.adam_step(grads, LR)
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts[1:]))
    return {"value": Bs / dt, "unit": "images/s", "cores": ncores, "kind": "port",
            "sample": "2 timed steps (1 warm-up) of %d images of the batch-64 workload, oracle torch-CPU fp32" % Bs}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--backbone", default="Xception", choices=["Xception", "MobileNet", "InceptionResNetV2"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.steps_ref = max(1, min(args.steps, 3))
    args.warmup_ref = 1
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    # keep stdout clean for the single JSON line (NCCL prints its version banner to stdout)
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    run_ours(args, rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
