#!/usr/bin/env python
"""bench.py — frames/sec of the north-star hot path (per-layer Conv2D / depthwise / pool / FC inference) on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (libsnn_b200.so, hand-written sm_100a CUDA)
  python bench.py --impl reference --gpus N ...            the reference's CPU operator path on the host cores

A "step" is one forward pass over one batch of synthetic images. Workload at N=1 = BASELINE.json configs[1]:
ResNet-18 (the reference's modelzoo/Resnet18 graph), 224x224x3, batch 32; with --gpus N every rank runs its own batch
of 32 (weak scaling, no collective on the forward path; the packed weight arena is broadcast once over NCCL at init).
`value` is device-timed with inputs resident in HBM (CUDA events on the engine's own stream, max over ranks); `e2e`
goes through the public C-ABI call snnb_model_run() with pinned HOST buffers (H2D of the batch + D2H of the logits
inside the timed region). One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (modelzoo key, batch per GPU, description)
    "resnet18": ("resnet18", 32, "ResNet-18 classification, 224x224x3, batch 32 per GPU (BASELINE.json configs[1])"),
    "mobilenetv2": ("mobilenetv2", 64, "MobileNetV2, 224x224x3, batch 64 per GPU (BASELINE.json configs[2], un-sharded)"),
    "yolov3tiny": ("yolov3tiny", 16, "YOLOv3-tiny, 416x416x3, batch 16 per GPU (BASELINE.json configs[3])"),
    "candy": ("candy", 8, "Fast-neural-style Candy, 720x720x3, batch 8 per GPU (BASELINE.json configs[4])"),
    "espcn": ("espcn", 1, "ESPCN 2x, 224x224x1, batch 1 (BASELINE.json configs[0])"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def layer_work(layers, shapes, batch):
    """Algorithmic work per layer: bytes = 4*(in + out + |W| + |b|) (+ 2nd input for Add),
    flops = 2*N*OH*OW*OC*(IC/groups)*k*k. `shapes[i]` = (N,H,W,C) of layer i's output as reported by the engine."""
    work = []
    for i, l in enumerate(layers):
        t = l["type"]
        if t == "InputLayer":
            work.append((t, 0.0, 0.0))
            continue
        ins = [shapes[j] for j in l.get("inputId", [])]
        n, oh, ow, oc = shapes[i]
        flops = 0.0
        wbytes = 0.0
        if t == "Conv2D":
            k, ic = l["kernel_size"], l["inputPlanes"]
            flops = 2.0 * n * oh * ow * oc * ic * k * k
            wbytes = 4.0 * (oc * ic * k * k + oc)
        elif t == "DepthwiseConv2D":
            k = l["kernel_size"]
            flops = 2.0 * n * oh * ow * oc * k * k
            wbytes = 4.0 * (oc * k * k + oc)
        elif t == "Dense":
            flops = 2.0 * n * l["units"] * l["inputPlanes"]
            wbytes = 4.0 * (l["units"] * l["inputPlanes"] + l["units"])
        inb = sum(4.0 * a * b * c * d for (a, b, c, d) in ins) if t != "YOLO" else 0.0
        outb = 4.0 * n * oh * ow * oc if t != "YOLO" else 0.0
        work.append((t, flops, inb + outb + wbytes))
    return work


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, region=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, sm_in, mx, reasons = [], [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                v0 = float(f[0])
                mx = float(f[1])
            except ValueError:
                continue
            sm.append(v0)
            # a sample is read a few ms after the driver took it: count it for the timed region if it arrived inside it or just after
            if region and region[0] <= ts <= region[1] + 0.02:
                sm_in.append(v0)
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        use = sm_in if sm_in else sm
        return {"sm_mhz": statistics.median(use) if use else None, "sm_max_mhz": mx, "samples": len(use), "samples_in_timed_region": len(sm_in),
                "samples_under_load": len(sm), "reasons": sorted(reasons)}


def tune_oracle_threads(oracle_mod, model, x1):
    """Pick the OpenMP team size that runs the oracle fastest on this host (all logical CPUs is often slower than the physical
    cores on an SMT box); torchrun exports OMP_NUM_THREADS=1, so the count is always set explicitly."""
    n = os.cpu_count() or 1
    best_t, best_dt = n, None
    for t in sorted({n, max(1, n // 2), max(1, n // 4)}, reverse=True):
        oracle_mod.lib().orc_set_num_threads(t)
        model.run(x1)
        t0 = time.perf_counter()
        model.run(x1)
        dt = time.perf_counter() - t0
        if best_dt is None or dt < best_dt:
            best_t, best_dt = t, dt
    oracle_mod.lib().orc_set_num_threads(best_t)
    return best_t


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU operator path. The reference has NO CPU Conv/Pool/BN, so the
    conv/pool/add body is the oracle's C++ restatement of its shader semantics ("port", OpenMP over all host cores) and
    the Dense/softmax tail goes through the reference's own compiled cpulayer.h when oracle/_ref is present."""
    if rank != 0:
        return 0
    from oracle import oracle
    from shadernn_b200 import modelzoo
    key, batch, desc = WORKLOADS[args.workload]
    d = tempfile.mkdtemp(prefix="snnb_bench_ref_")
    path, layers = modelzoo.build(key, d)
    hw = modelzoo.MODELS[key][1]
    sample = max(1, min(batch, args.cpu_sample))
    x = modelzoo.synthetic_input(key, sample)
    m = oracle.Model(path)
    tune_oracle_threads(oracle, m, x[:1])
    threads = oracle.lib().orc_num_threads()
    for _ in range(max(1, min(args.warmup, 2))):
        m.run(x)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        m.run(x)
    dt = time.perf_counter() - t0
    fps = sample * steps / dt
    line = {
        "impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2),
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "input_hw": list(hw), "sample_frames_per_step": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frames per step x %d steps of the same graph/weights; oracle C++ restatement (OpenMP, %d threads)" % (sample, steps, threads)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# (mean4, norm4) of ImageTexture::convertToRGBA32FAndNormalize per model (demo/common/modelInference.cpp:135-224)
U8_NORM = {
    "resnet18": ([127.5] * 4, [1.0 / 127.5] * 4),
    "yolov3tiny": ([127.5] * 4, [1.0 / 127.5] * 4),
    "mobilenetv2": ([0.0] * 4, [1.0 / 255.0] * 4),
    "candy": ([0.0] * 4, [1.0] * 4),
    "espcn": ([0.0] * 4, [1.0 / 255.0] * 4),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="resnet18", choices=sorted(WORKLOADS))
    ap.add_argument("--algo", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-fuse", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=4, help="frames per step of the CPU baseline / reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", action="store_true", help="also print the per-layer roofline table to stderr")
    ap.add_argument("--precision", default="fp32x3", choices=["fp32x3", "fp16w", "fp16"],
                    help="product form of the tensor-core path (snnb.h SNNB_PRECISION_*); the headline is fp32x3")
    ap.add_argument("--batch", type=int, default=0, help="override the workload's batch per GPU (the metric's config is the default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from shadernn_b200 import parallel
    rank, local_rank, world = parallel.env_world()
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    from shadernn_b200 import core, modelzoo
    from shadernn_b200._lib import lib, check
    if world > 1:
        parallel.init_distributed("nccl")
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank

    key, batch, desc = WORKLOADS[args.workload]
    if args.batch > 0 and args.batch != batch:
        batch, desc = args.batch, desc + " [batch overridden to %d]" % args.batch
    hw = modelzoo.MODELS[key][1]
    d = tempfile.mkdtemp(prefix="snnb_bench_r%d_" % rank)
    path, layers = modelzoo.build(key, d)
    ctx = core.GpuContext(local_rank)
    model = core.MixedInferenceCore(ctx, path, batch=batch, conv_algo=args.algo, use_cuda_graph=not args.no_graph, fuse=not args.no_fuse, precision=args.precision)
    arena_bytes = parallel.broadcast_model_weights(model, dev, src=0) if world > 1 else model.weight_arena()[1]

    x = modelzoo.synthetic_input(key, batch, seed=7767517 + rank)
    in_shape, out_shape = model.input_shape(0), model.output_shape(0)
    host_in = torch.from_numpy(x).pin_memory()
    host_out = torch.empty(int(np.prod(out_shape)), dtype=torch.float32).pin_memory()
    classes = torch.zeros(batch, dtype=torch.int32).pin_memory()

    # ---- device-resident throughput ("value") ----
    model.set_input(x)
    for _ in range(args.warmup):
        model.forward()
    ctx.sync()
    tm = C.c_void_p()
    check(lib().snnb_timer_create(ctx.h, C.byref(tm)))
    sampler = ClockSampler(local_rank)
    sampler.start()
    # keep the GPU under the same load while nvidia-smi starts up (~0.1 s), so that its 10 ms samples fall inside the timed
    # region and every sample it ever takes is a sample under load
    t_load = time.perf_counter()
    while not sampler.lines and time.perf_counter() - t_load < 3.0:
        for _ in range(8):
            model.forward()
        ctx.sync()
    parallel.barrier()
    ctx.sync()
    launches0 = ctx.launches
    t_region0 = time.perf_counter()
    check(lib().snnb_timer_start(tm))
    for _ in range(args.steps):
        model.forward()
    check(lib().snnb_timer_stop(tm))
    ms = C.c_float()
    check(lib().snnb_timer_elapsed_ms(tm, C.byref(ms)))
    ctx.sync()
    t_region1 = time.perf_counter()
    launches = ctx.launches - launches0
    clocks = sampler.stop((t_region0, t_region1))
    parallel.barrier()
    dev_ms = parallel.max_over_ranks(ms.value, dev)

    # ---- end to end through the C-ABI with host buffers ("e2e") ----
    # The serving loop a user writes: snnb_model_submit / snnb_model_wait (double-buffered: the H2D copy of batch i+1
    # overlaps the forward pass of batch i). EVERY step uploads its own fp32 batch from pinned host memory and downloads
    # its logits + class indices; the timed region is host wall-clock around K such steps between device syncs.
    host_in2 = torch.from_numpy(modelzoo.synthetic_input(key, batch, seed=99 + rank)).pin_memory()
    host_out2 = torch.empty_like(host_out).pin_memory()
    classes2 = torch.zeros_like(classes).pin_memory()
    bufs = [(host_in, host_out, classes), (host_in2, host_out2, classes2)]

    streaming = model.num_outputs >= 1 and all(l["type"] != "YOLO" for l in layers)  # detection decodes on the host: synchronous run()

    def e2e_loop(steps):
        if not streaming:
            for i in range(steps):
                hin, hout, hcls = bufs[i & 1]
                model.run_raw(hin.data_ptr(), hout.data_ptr(), hout.numel(), hcls.data_ptr())
            return
        pending = None
        for i in range(steps):
            hin, hout, hcls = bufs[i & 1]
            t = model.submit_raw(hin.data_ptr(), hout.data_ptr(), hout.numel(), hcls.data_ptr())
            if pending is not None:
                model.wait(pending)
            pending = t
        model.wait(pending)

    e2e_loop(4)
    parallel.barrier()
    ctx.sync()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    ctx.sync()
    e2e_ms = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
    parallel.barrier()

    # The same loop fed with 8-bit images (what the reference's demo apps start from): snnb_model_submit_u8 normalises on the
    # device as ImageTexture::convertToRGBA32FAndNormalize does (imageTexture.h:114; constants of modelInference.cpp), so a
    # quarter of the bytes cross PCIe. Every step still uploads its own batch and downloads logits + class indices.
    u8_ms = None
    if streaming:
        mean4, norm4 = U8_NORM[key]
        rng8 = np.random.default_rng(1234 + rank)
        u8bufs = [torch.from_numpy(rng8.integers(0, 256, in_shape, dtype=np.uint8)).pin_memory() for _ in range(2)]

        def u8_loop(steps):
            pending = None
            for i in range(steps):
                _, hout, hcls = bufs[i & 1]
                t = model.submit_u8_raw(u8bufs[i & 1].data_ptr(), mean4, norm4, hout.data_ptr(), hout.numel(), hcls.data_ptr())
                if pending is not None:
                    model.wait(pending)
                pending = t
            model.wait(pending)

        u8_loop(4)
        parallel.barrier()
        ctx.sync()
        t0 = time.perf_counter()
        u8_loop(args.steps)
        ctx.sync()
        u8_ms = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
        parallel.barrier()
    # the strictly synchronous call (one batch at a time, nothing overlapped), for reference
    for _ in range(2):
        model.run_raw(host_in.data_ptr(), host_out.data_ptr(), host_out.numel(), classes.data_ptr())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.run_raw(host_in.data_ptr(), host_out.data_ptr(), host_out.numel(), classes.data_ptr())
    ctx.sync()
    sync_ms = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
    parallel.barrier()

    if world > 1:
        import torch.distributed as dist
        parallel.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    # ---- roofline of the dominant kernel (per-layer event pairs, eager pass, live in this process) ----
    pk = peaks()
    lt = np.zeros(model.num_layers, np.float64)
    reps = 5
    model.time_layers()
    for _ in range(reps):
        lt += model.time_layers()
    lt /= reps
    work = layer_work(layers, [model.layer_info(i)[2] for i in range(model.num_layers)], batch)
    by_kind = {}
    for (t, fl, by), ms_l in zip(work, lt):
        k = by_kind.setdefault(t, [0.0, 0.0, 0.0, 0])
        k[0] += ms_l
        k[1] += fl
        k[2] += by
        k[3] += 1 if ms_l > 0 else 0
    dom = max(by_kind, key=lambda t: by_kind[t][0])
    dms, dfl, dby, dn = by_kind[dom]
    tmin_tensor = dfl / (pk["bf16_tflops_sustained"] * 1e12) if dfl else 0.0
    tmin_hbm = dby / (pk["hbm_gbs"] * 1e9)
    if dfl and tmin_tensor >= tmin_hbm:
        roof = {"bound": "tensor", "achieved": dfl / (dms * 1e-3) / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s"}
    else:
        roof = {"bound": "hbm", "achieved": dby / (dms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = None
    # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of this command (profiles/README.md),
    # scaled to the launches of one step like `algorithmic_per_step`: far BELOW the algorithmic bytes here because a layer's
    # input is still L2-resident from its producer (126 MB L2) - the re-reads that matter are L2->SM, see l2_to_sm_read_MB.
    prof = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_%s_kernels.csv" % args.workload)
    if dom == "Conv2D" and os.path.exists(prof):
        import csv
        rows = list(csv.reader(open(prof)))
        col = {n.split("[")[0]: i for i, n in enumerate(rows[0])}
        sel = [r for r in rows[1:] if "conv_umma" in r[col["kernel"]] and not r[col["grid"]].startswith("(1,")]
        if sel:
            mb = [float(r[col["dram_read_MB"]]) + float(r[col["dram_write_MB"]]) * (1e-3 if "Kbyte" in rows[0][col["dram_write_MB"]] else 1.0) for r in sel]
            l2 = [float(r[col["l2_to_sm_read_MB"]]) for r in sel]
            roof["traffic"] = sum(mb) / len(mb) * 1e6 * dn
            roof["traffic_note"] = "mean DRAM read+write of %d profiled conv_umma launches x %d launches/step (%s); L2->SM reads %.0f MB/launch" % (
                len(sel), dn, os.path.basename(prof), sum(l2) / len(l2))
    if dfl:  # each fp32-equivalent product is executed as 3 bf16 MMAs (split-bf16): the tensor pipe's own view of the same layers
        roof["tensor_executed"] = {"flops_per_step": 3.0 * dfl, "achieved_tflops": 3.0 * dfl / (dms * 1e-3) / 1e12,
                                   "frac_of_bf16_peak": 3.0 * dfl / (dms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"]}
    roof["kernel"] = "%s layers (%d launches/step, %.3f ms/step of %.3f ms eager total)" % (dom, dn, dms, float(lt.sum()))
    roof["peak_source"] = pk["source"] + ("; sustained bf16 dense (kernel timed inside a long step)" if roof["bound"] == "tensor" else "")
    roof["algorithmic_per_step"] = {"flops": dfl, "bytes": dby}
    # whole-graph lower bound: sum over layers of max(bytes/BW, flops/peak)
    tmin = sum(max(by / (pk["hbm_gbs"] * 1e9), fl / (pk["bf16_tflops_sustained"] * 1e12)) for (_, fl, by) in work)
    roof["graph_frac"] = tmin / (dev_ms / args.steps * 1e-3)
    if args.layers:
        sys.stderr.write("# %s, batch %d per GPU; per-layer event pairs (eager pass, %d repetitions). roofline%% = max(bytes/HBM, flops/bf16) / t with the\n"
                         "# algorithmic work of SURVEY 8d (fp32 bytes, 2*MAC flops); tensor_exec%% = 3 x flops / bf16 peak / t (each product is three bf16 MMAs)\n"
                         % (desc, batch, reps))
        sys.stderr.write("# %-3s %-22s %9s %9s %9s %9s %9s %6s %10s %12s\n" % ("id", "layer", "ms", "GFLOP", "MB", "TF/s", "GB/s", "bound", "roofline%", "tensor_exec%"))
        for i, ((t, fl, by), ms_l) in enumerate(zip(work, lt)):
            if ms_l <= 0:
                continue
            t_h, t_t = by / (pk["hbm_gbs"] * 1e9), fl / (pk["bf16_tflops_sustained"] * 1e12)
            sys.stderr.write("[%02d] %-22s %9.3f %9.2f %9.2f %9.1f %9.0f %6s %9.1f%% %11.1f%%\n" %
                             (i, t, ms_l, fl / 1e9, by / 1e6, fl / ms_l / 1e9, by / ms_l / 1e6, "tensor" if t_t > t_h else "hbm",
                              100 * max(t_h, t_t) / (ms_l * 1e-3), 100 * 3 * t_t / (ms_l * 1e-3)))
        sys.stderr.write("# total eager %.3f ms; CUDA-graph step %.3f ms\n" % (float(lt.sum()), dev_ms / args.steps))

    # ---- CPU baseline beside it (bounded sample, rank 0, N=1 only) ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle
        sample = max(1, min(batch, args.cpu_sample))
        om = oracle.Model(path)
        xs = x[:sample]
        tune_oracle_threads(oracle, om, x[:1])
        om.run(xs)
        t0 = time.perf_counter()
        reps_c = 3
        for _ in range(reps_c):
            ref_out = om.run(xs)
        dtc = (time.perf_counter() - t0) / reps_c
        threads = oracle.lib().orc_num_threads()
        cpu = {"value": sample / dtc, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "%d frames x %d reps of the same graph, weights and inputs; oracle C++ restatement of the reference operators, OpenMP %d threads" %
                         (sample, reps_c, threads)}
        # the bench doubles as a parity spot-check on the full-size graph
        got = host_out.numpy().reshape(out_shape)[:sample]
        if ref_out.shape == got.shape:
            cpu["parity_mismatches_vs_oracle"] = int(oracle.compare(got, ref_out, 1e-3))

    frames = batch * world * args.steps
    line = {
        "metric": "frames/sec", "value": frames / (dev_ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (split-bf16 hi+lo storage, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": desc, "model": key, "input_hw": list(hw), "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": "dp%d" % world,
                   "conv_algo": args.algo, "cuda_graph": not args.no_graph, "fused": not args.no_fuse, "weights_broadcast_bytes": arena_bytes,
                   "l2": "per-step working set (~%.1f GB of activations) exceeds the 126 MB L2; no explicit flush" % (sum(b for _, _, b in work) / 1e9)},
        "clocks": clocks,
        "e2e": ({"value": frames / (u8_ms * 1e-3), "unit": "frames/s", "ms_per_step": u8_ms / args.steps,
                 "h2d_bytes_per_step": int(np.prod(in_shape)), "d2h_bytes_per_step": int(np.prod(out_shape)) * 4 + batch * 4,
                 "api": "snnb_model_submit_u8/snnb_model_wait (double-buffered; pinned host 8-bit NHWC images in, normalised on the device as the "
                        "reference's ImageTexture does; logits + class indices out)",
                 "fp32_input": {"value": frames / (e2e_ms * 1e-3), "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(np.prod(in_shape)) * 4,
                                "api": "snnb_model_submit/snnb_model_wait with pinned host fp32 NHWC (PCIe-bound: 4x the bytes)"},
                 "synchronous_run_frames_per_s": frames / (sync_ms * 1e-3)} if u8_ms else
                {"value": frames / (e2e_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_ms / args.steps,
                 "h2d_bytes_per_step": int(np.prod(in_shape)) * 4, "d2h_bytes_per_step": int(np.prod(out_shape)) * 4 + batch * 4,
                 "api": "snnb_model_run (synchronous, pinned host fp32 NHWC in; YOLO decode + NMS on the host inside the timed region)",
                 "synchronous_run_frames_per_s": frames / (sync_ms * 1e-3)}),
        "gpu_launches": int(launches),
        "roofline": roof,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
