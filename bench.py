#!/usr/bin/env python
"""bench.py -- throughput of the per-step training hot path (BASELINE.json: "3D patches/sec (EWC on)").

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (N>1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's PyTorch path on the host CPU cores

Workload (config.workload, default cfg2): BASELINE.json configs[1] -- nnUNetTrainerEWC, 5-stage 3D U-Net, 64x128x128
synthetic hippocampus-shaped patches, batch 2 per GPU, one stored EWC task, bf16 activations (fp32 accumulation, fp32
params).  A "step" = zero_grad -> forward -> Dice+CE (+EWC) -> backward -> [grad all-reduce] -> clip(12) -> SGD-Nesterov.
--workload cfg3 | cfg4 | cfg5 run the other BASELINE.json configurations (parity-test cases, benched for the record):
  cfg3  nnUNetTrainerLWF, 64x160x160, one finished task (its head evaluated on the shared body + KL term)
  cfg4  nnUNetTrainerPLOP + Generic_ViT_UNet (V1, base), 48x192x192, teacher in the loop (pseudo labels + local POD)
  cfg5  nnUNetTrainerRW, 64x128x128, third task of a sequence (two stored tasks penalised: documented math,
        strict_reference=False -- under the reference's quirk Q2 the penalty would be off after the first iteration),
        Fisher / score update every 10 iterations

  value        patches/s with the batch already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e          patches/s through trainer.run_iteration(generator) with pinned HOST buffers: H2D of data+targets and
               D2H of the loss inside the timed region
  roofline     the dominant kernel (3x3x3 conv forward of the full-resolution 32->32 layer), timed live with CUDA
               events: achieved = algorithmic FLOPs / duration, peak = MEASURED_PEAKS.json bf16 burst
  cpu_baseline the oracle port of the reference step (PyTorch CPU fp32) on a bounded sample, rank 0, N=1 only
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "lifelong-nnunet_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

# keep stdout to the one JSON line: NCCL prints its version banner there when NCCL_DEBUG is VERSION / unset in some images
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402

WORKLOAD = "cfg2"
METRIC = {"cfg1": "3D patches/sec", "cfg2": "3D patches/sec (EWC on)", "cfg3": "3D patches/sec (LwF on)",
          "cfg4": "3D patches/sec (PLOP + ViT)", "cfg5": "3D patches/sec (RW on)"}
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["source"] = "measured"
        return d
    return dict(FALLBACK_PEAKS)


TRAINER_DESC = {
    "cfg1": "nnUNetTrainerSequential",
    "cfg2": "nnUNetTrainerEWC, EWC on (1 stored task, lambda 0.4)",
    "cfg3": "nnUNetTrainerLWF, 1 finished task (old head on the shared body + KL, T=2)",
    "cfg4": "nnUNetTrainerPLOP + Generic_ViT_UNet V1 base (native ViT: tcgen05 GEMMs + warp-MMA attention), frozen teacher in the loop, pod_lambda 1e-2, 3 scales",
    "cfg5": "nnUNetTrainerRW, task 3 of 3 (2 stored tasks penalised every iteration: strict_reference=False), F/S update every 10 its",
}


def workload_desc(geom, precision, n_gpus):
    return {"workload": "%s: %s, %d-stage 3D U-Net, %dx%dx%d synthetic 1-ch patches, batch %d/GPU, %s activations" %
                        (geom.name, TRAINER_DESC.get(geom.name, "nnUNetTrainerEWC"), geom.num_pool, geom.patch[0], geom.patch[1],
                         geom.patch[2], geom.batch, precision),
            "global_batch": geom.batch * n_gpus, "patch": list(geom.patch), "parallelism": "dp%d" % n_gpus,
            "pool_op_kernel_sizes": [list(k) for k in geom.pool], "num_classes": geom.num_classes,
            "l2_policy": "inputs+activations (>3 GB per step) exceed the 126 MB L2; no explicit flush",
            "fwd_bwd_gflop_per_patch": round(3 * geom.fwd_flops_per_patch() / 1e9, 1)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm: the oracle port of the reference's PyTorch step on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(geom, batch):
    """the oracle port of the reference's iteration for the workload's trainer (PyTorch CPU fp32, all host threads)"""
    import copy
    import numpy as np
    from b200unet import synth
    from oracle import cl_losses, step
    torch.set_num_threads(os.cpu_count() or 1)
    weights = cl_losses.ds_loss_weights(geom.num_pool)
    data, targets = synth.make_batch(geom, batch=batch)
    pools = [list(k) for k in geom.pool]
    if geom.name == "cfg4":
        from oracle import vit_unet
        torch.manual_seed(0)
        net = vit_unet.Generic_ViT_UNet(geom.in_channels, geom.base_features, geom.num_classes, geom.num_pool,
                                        [int(s) for s in geom.patch], pools)
    else:
        net = step.build_network(geom.in_channels, geom.base_features, geom.num_classes, pools, max_num_features=geom.max_features)
    opt = step.make_optimizer(net)
    if geom.name == "cfg3":       # LwF: one extra (no-grad) forward per old head + KL against the stored logits (lwf:298-370)
        old_head = copy.deepcopy(net.seg_outputs.state_dict())
        with torch.no_grad():
            stored = net(data)[0]
        base = step.base_loss_fn(weights)

        def one():
            cur = copy.deepcopy(net.seg_outputs.state_dict())
            net.seg_outputs.load_state_dict(old_head)
            with torch.no_grad():
                pred = net(data)[0]
            net.seg_outputs.load_state_dict(cur)
            return step.run_iteration(net, opt, data, targets, lambda o, t: base(o, t) + cl_losses.lwf_distillation(pred, stored, 2.0))[0]
        return one
    if geom.name == "cfg4":       # PLOP: teacher forward, hooks on every conv of both nets, pseudo labels + local POD (plop:217-328)
        teacher = copy.deepcopy(net)
        acts, acts_o = {}, {}
        for name, m in net.named_modules():
            if 'conv.Conv' in str(type(m)):
                m.register_forward_hook(lambda mod, i, o, name=name: acts.__setitem__(name, o.detach()))
        for name, m in teacher.named_modules():
            if 'conv.Conv' in str(type(m)):
                m.register_forward_hook(lambda mod, i, o, name=name: acts_o.__setitem__(name, o.detach()))
        thr = {i: torch.full((geom.num_classes,), 1e-3) for i in range(geom.num_pool)}

        def loss_fn(out, tgt):
            with torch.no_grad():
                out_o = teacher(data)
            a = {k: v for k, v in acts.items() if v.dim() == 5}
            b = {k: v for k, v in acts_o.items() if v.dim() == 5}
            return cl_losses.plop_loss(out, out_o, tgt, weights, thr, float(np.log(geom.num_classes)), a, b, 1e-2, 3)
        return lambda: step.run_iteration(net, opt, data, targets, loss_fn)[0]
    if geom.name == "cfg5":       # RW: two stored tasks, penalty every iteration, F / S update every 10 iterations (rw:231-265)
        named = list(net.named_parameters())
        fisher, params, scores = {}, {}, {}
        for i, t in enumerate(("A", "B")):
            fisher[t], params[t], scores[t] = synth.make_ewc_state(named, seed=7 + i, with_scores=True)
        fisher["C"] = {n: torch.zeros_like(p) for n, p in named}
        sc = {n: torch.zeros_like(p) for n, p in named}
        state = {"count": 0, "prev": None}
        loss_fn = step.rw_loss_fn(net, weights, fisher, params, scores, 0.4, strict_reference=False)

        def one():
            l = step.run_iteration(net, opt, data, targets, loss_fn)[0]
            if state["count"] % 10 == 0:
                prev = state["prev"]
                for n, p in named:
                    if p.grad is not None:
                        fisher["C"][n], sc[n] = cl_losses.rw_update(p.detach(), p.grad.detach(), None if prev is None else prev[n],
                                                                    fisher["C"][n], sc[n], 0.9)
                state["prev"] = {n: p.detach().clone() for n, p in named}
            state["count"] += 1
            return l
        return one
    if geom.name == "cfg1":
        loss_fn = step.base_loss_fn(weights)
    else:
        fisher, params = synth.make_ewc_state(list(net.named_parameters()))
        loss_fn = step.ewc_loss_fn(net, weights, {"A": fisher}, {"A": params}, 0.4)
    return lambda: step.run_iteration(net, opt, data, targets, loss_fn)[0]


def time_cpu(geom, batch, steps, warmup, budget_s=None):
    """per-step wall times of the CPU arm; with `budget_s` the number of timed steps is cut so that the whole run stays
    within about that many seconds (a CPU step of the full workload takes seconds to tens of seconds)"""
    one = cpu_reference_step_fn(geom, batch)
    t_first = None
    for _ in range(warmup):
        t0 = time.perf_counter()
        one()
        t_first = time.perf_counter() - t0
    ts = []
    for i in range(steps):
        t0 = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t0)
        est = t_first if t_first is not None else ts[0]
        if budget_s is not None and (warmup + len(ts) + 1) * est > budget_s:
            break
    return ts


def run_reference(args, geom):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, min(args.warmup, 1)   # bounded: every CPU step is seconds long
    ts = time_cpu(geom, geom.batch, steps, warmup, budget_s=150.0)    # bounded sample: the run ends within a few minutes
    total = sum(ts)
    val = geom.batch * len(ts) / total
    sample = "%d timed + %d warm-up steps of the full workload batch (B=%d), oracle port (PyTorch CPU fp32, eager)" % (len(ts), warmup, geom.batch)
    line = {"impl": "reference", "metric": METRIC.get(geom.name, METRIC["cfg2"]), "value": val, "unit": "patches/s", "n_gpus": args.gpus,
            "steps": len(ts), "warmup": warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_desc(geom, "fp32 (CPU)", 1),
            "cpu_baseline": {"value": val, "unit": "patches/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an NVML polling thread (every 5 ms; the timed region of
    the default run is ~100 ms, shorter than nvidia-smi's start-up), `nvidia-smi -lms` as the fallback."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc, self.index = None, index
        self.thread, self.stop_flag, self.samples, self.reasons, self.mx = None, False, [], set(), None

    def _poll(self, nv, h):
        masks = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, m in masks.items():
                    if r & m:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = self.samples
            hi = sorted(sm)[len(sm) // 2:] if sm else []
            return {"sm_mhz": statistics.median(hi) if hi else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml, 5 ms period, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        hi = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(hi) if hi else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


def pinned_batch_generator(data, targets):
    hd = data.pin_memory()
    ht = [t.pin_memory() for t in targets]
    while True:
        yield {'data': hd, 'target': ht}


def device_batch_generator(data, targets):
    while True:
        yield {'data': data, 'target': targets}


def time_dominant_kernel(geom, precision, steps, warmup):
    """CUDA-event timing of the dominant kernel in isolation: forward 3x3x3 conv of the full-resolution
    base->base layer (conv_blocks_context.0.blocks.1) at the workload's batch, exactly as the training step launches it
    (one kernel: convolution + bias + InstanceNorm partial sums in the epilogue)."""
    import ctypes as C
    from b200unet import _lib
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    B, (D, H, W), Cc = geom.batch, geom.patch, geom.base_features
    dt = torch.float32 if precision == "fp32" else torch.bfloat16
    x = torch.randn((B, D, H, W, Cc), device=dev).to(dt)
    z = torch.empty_like(x)
    w = torch.randn((Cc, Cc, 3, 3, 3), device=dev) * 0.05
    bias = torch.zeros(Cc, device=dev)
    stats = torch.empty((B, Cc, 2), device=dev)
    desc = _lib.ConvDesc()
    desc.n, desc.d, desc.h, desc.w, desc.cin, desc.cout = B, D, H, W, Cc, Cc
    for i in range(3):
        desc.stride[i] = 1
    desc.in_pitch = desc.out_pitch = Cc
    desc.dtype = _lib.B2_F32 if precision == "fp32" else _lib.B2_BF16
    scr = torch.empty(int(lib.b2_conv3d_scratch_bytes(C.byref(desc))), dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    if precision == "bf16":
        # prepared weights: the timed call launches exactly ONE kernel (the convolution)
        shadow = torch.empty(int(lib.b2_conv3d_shadow_bytes(C.byref(desc))), dtype=torch.uint8, device=dev)
        _lib.check(lib.b2_conv3d_make_shadow(C.byref(desc), w.data_ptr(), shadow.data_ptr(), st))

        def call():    # the PRODUCT variant: InstanceNorm partial sums from the epilogue, as the training step launches it
            _lib.check(lib.b2_conv3d_fwd_shadow_stats(C.byref(desc), x.data_ptr(), shadow.data_ptr(), bias.data_ptr(), z.data_ptr(),
                                                      scr.data_ptr(), st))
    else:
        def call():
            _lib.check(lib.b2_conv3d_fwd(C.byref(desc), x.data_ptr(), w.data_ptr(), bias.data_ptr(), z.data_ptr(),
                                         stats.data_ptr(), 1e-5, scr.data_ptr(), st))
    for _ in range(max(warmup, 3)):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(steps, 5)
    e0.record()
    for _ in range(n):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * B * D * H * W * Cc * Cc * 27
    return ms, flops


def dominant_traffic(geom):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/dominant_kernel.json); null for workloads other than the profiled one."""
    p = os.path.join(ROOT, "profiles", "dominant_kernel.json")
    if geom.name not in ("cfg2", "cfg5") or not os.path.exists(p):
        return None
    d = json.load(open(p))
    return {"value": d["dram_bytes_read"] + d["dram_bytes_write"], "unit": "bytes/launch",
            "algorithmic_bytes": d["algorithmic_bytes"], "source": d["source"]}


def build_trainer(geom, precision, dev, ddp, cuda_graph):
    """the trainer of the workload with its continual-learning state (synthetic)"""
    from b200unet import synth
    from b200unet import trainers as T
    kw = dict(precision=precision, device=dev, ddp=ddp, seed=0, cuda_graph=cuda_graph)
    if geom.name == "cfg3":
        tr = T.nnUNetTrainerLWF(geom, task="task_prev", **kw)
        tr.initialize()
        tr.finish_task()
        tr.start_task("task_new")
        data, _ = synth.make_batch(geom, seed=99)
        tr.store_target_logits([data])
        return tr
    if geom.name == "cfg4":
        tr = T.nnUNetTrainerPLOP(geom, use_vit=True, **kw)
        tr.initialize()
        tr.start_new_task()
        return tr
    if geom.name == "cfg5":
        tr = T.nnUNetTrainerRW(geom, strict_reference=False, fisher_update_after=10, **kw)
        tr.initialize()
        named = list(tr.network.named_parameters())
        for i, t in enumerate(("task_A", "task_B")):
            f, p, s_ = synth.make_ewc_state(named, seed=7 + i, with_scores=True)
            tr.fisher[t] = {k: v.to(dev) for k, v in f.items()}
            tr.params[t] = {k: v.to(dev) for k, v in p.items()}
            tr.scores[t] = {k: v.to(dev) for k, v in s_.items()}
        tr.start_task("task_C")
        return tr
    if geom.name == "cfg1":
        tr = T.nnUNetTrainerSequential(geom, **kw)
        tr.initialize()
        return tr
    tr = T.nnUNetTrainerEWC(geom, **kw)
    tr.initialize()
    fisher, params = synth.make_ewc_state(list(tr.network.named_parameters()), seed=7)
    tr.fisher["task_prev"] = {k: v.to(dev) for k, v in fisher.items()}
    tr.params["task_prev"] = {k: v.to(dev) for k, v in params.items()}
    tr.loss.update_ewc_params(tr.fisher, tr.params)
    tr.loss.update_network_params(tr.network.named_parameters())
    return tr


def in_step_kernel_us():
    """duration of the dominant kernel INSIDE a training step, from the committed ncu launch list of this round"""
    p = os.path.join(ROOT, "profiles", "dominant_kernel.json")
    if os.path.exists(p):
        return json.load(open(p)).get("in_step_us")
    return None


def run_ours(args, geom):
    import torch.distributed as dist
    from b200unet import _lib, synth
    from b200unet.trainers import DataParallelGroup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ddp = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ddp = DataParallelGroup()
    precision = args.precision
    trainer = build_trainer(geom, precision, dev, ddp, cuda_graph=not args.no_graph)

    data, targets = synth.make_batch(geom, seed=1234 + rank)     # rank-seeded patches
    h2d = data.numel() * 4 + sum(t.numel() * 4 for t in targets)
    gen_host = pinned_batch_generator(data, targets)
    gen_dev = device_batch_generator(data.to(dev), [t.to(dev) for t in targets])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(gen, steps, detach):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            trainer.run_iteration(gen, do_backprop=True, run_online_evaluation=False, detach=detach)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    # launches per step: counted on an eager (not yet captured) iteration; a captured step replays exactly these
    trainer.run_iteration(gen_dev, detach=False)
    l0 = _lib.launch_count()
    trainer.run_iteration(gen_dev, detach=False)
    launches_per_step = _lib.launch_count() - l0
    warm = max(args.warmup, 3)
    for _ in range(warm):
        trainer.run_iteration(gen_dev, detach=False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, _ = timed(gen_dev, args.steps, detach=False)
    clocks = sampler.stop() if rank == 0 else None
    # e2e, default path: pinned HOST batches through trainer.run_iteration, every step copies its batch to the device (input
    # patch on the compute stream, targets on a copy stream next to the forward pass) and reads its loss back to the host
    for _ in range(2):
        trainer.run_iteration(gen_host, detach=True)
    _, ms_e2e = timed(gen_host, args.steps, detach=True)
    # e2e with the opt-in input pipelining (the H2D copy of batch i+1 overlaps step i; draws the generator one batch ahead)
    trainer.prefetch_inputs = True
    for _ in range(2):
        trainer.run_iteration(gen_host, detach=True)
    _, ms_e2e_pf = timed(gen_host, args.steps, detach=True)
    trainer.prefetch_inputs = False

    patches = geom.batch * world * args.steps
    value = patches / (ms_dev / 1e3)
    e2e = patches / (ms_e2e / 1e3)
    e2e_pf = patches / (ms_e2e_pf / 1e3)
    if rank != 0:
        finish(world)
        return
    peaks = load_peaks()
    gflop_patch = 3 * geom.fwd_flops_per_patch() / 1e9
    kms, kflops = time_dominant_kernel(geom, precision, args.steps, warm)
    peak_tf = peaks["bf16_tflops"]
    achieved_tf = kflops / (kms * 1e-3) / 1e12
    step_tf = value / world * gflop_patch / 1e3
    graphed = any(getattr(s_, "graph", None) is not None for s_ in trainer._steps.values())
    line = {"metric": METRIC.get(geom.name, METRIC["cfg2"]), "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_desc(geom, precision, world),
            "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "api": "trainer.run_iteration(generator of pinned host batches), default path (no look-ahead on the generator)",
                    "with_prefetch_inputs": {"value": e2e_pf, "ms_per_step": ms_e2e_pf / args.steps,
                                             "note": "opt-in: H2D of batch i+1 overlaps step i (draws the generator one batch ahead)"}},
            "gpu_launches": int(launches_per_step * args.steps),
            "launches": {"per_step": int(launches_per_step), "cuda_graph": bool(graphed),
                         "note": "kernels of libb2unet per step (counted on an eager iteration); with cuda_graph the timed steps "
                                 "replay exactly these launches from two captured graphs (forward | loss+backward+optimiser)"},
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv3d 3x3x3 forward + InstanceNorm partial sums (stats epilogue ON, the variant the "
                                  "step runs), %d->%d @ %dx%dx%d x B%d (conv_blocks_context.0.blocks.1)" %
                                  (geom.base_features, geom.base_features, geom.patch[0], geom.patch[1], geom.patch[2], geom.batch),
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                         "peak_source": peaks["source"] + " bf16 burst (cuBLAS 8192^3)", "kernel_ms": kms,
                         "in_step_kernel_us": in_step_kernel_us(),
                         "traffic": dominant_traffic(geom),
                         "step": {"achieved": step_tf, "unit": "TFLOP/s", "frac_of_burst": step_tf / peak_tf,
                                  "frac_of_sustained": step_tf / peaks.get("bf16_tflops_sustained", peak_tf),
                                  "flops": "conv-stack fwd+bwd %.1f GFLOP/patch (teacher / ViT FLOPs not counted)" % gflop_patch},
                         "step_frac_of_sustained_peak": step_tf / peaks.get("bf16_tflops_sustained", peak_tf)}}
    if world == 1 and not args.no_cpu_baseline:
        ts = time_cpu(geom, geom.batch, 1, 1)
        line["cpu_baseline"] = {"value": geom.batch / ts[0], "unit": "patches/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "1 timed step (after 1 warm-up) of the full workload batch (B=%d), oracle port of the "
                                          "reference step (PyTorch CPU fp32 eager, %d threads)" % (geom.batch, os.cpu_count())}
    print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    """leave without tearing NCCL down: destroying a process group whose collectives were captured in CUDA graphs can
    block at exit; every rank has produced its result by now, so synchronise and exit hard"""
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every step instead of replaying the captured CUDA graphs")
    args = ap.parse_args()
    from b200unet.configs import CONFIGS
    geom = CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, geom)
    else:
        run_ours(args, geom)


if __name__ == "__main__":
    main()
