"""Benchmark of the Euler-integration hot path on B200 (contract: see the task prompt / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs-per-step P]

Headline workload (BASELINE.json configs[1]): reflow pair generation -- 100-step Euler through the default UNet
(64x64x3, 11.26 M parameters, seeded random-init weights because the reference's checkpoints are not in its
checkout) over host-seeded N(0,1) noise, sharded along the batch over the ranks.  One "step" = one batch of
P pairs per GPU integrated for 100 Euler steps.  metric = pairs/s, whole job.

Also reported on rank 0 (N=1): Euler sampling images/s at 1/2/4/8 steps for batch 64 (configs[0]) and batch 4096
(configs[2]), the per-kernel-class time split, the tcgen05 conv roofline, and the CPU baseline (oracle port timed
on the host cores on a bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAGE, CH, EULER_STEPS, TOTAL_PAIRS = 64, 3, 100, 65536
FLOPS_PER_IMG_STEP = 12.7636e9  # SURVEY.md §8d / BASELINE.md §3 (algorithmic, 64x64)
# DRAM bytes (read + write) per image of each kernel class in one velocity evaluation, written by tools/ncu_traffic.py from an
# `ncu --set full` capture of the CURRENT build (the file records the library digest it was captured with).
NCU_TRAFFIC_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_ncu_traffic.json")
GN_ELEMS_PER_IMAGE = 5013504    # elements normalised per velocity evaluation (30 GroupNorm sites)


def ncu_traffic(kind: str, micro_batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one kernel class per forward, from the committed ncu capture; None when
    the capture is missing or was taken with a different build of the library."""
    try:
        with open(NCU_TRAFFIC_FILE) as f:
            d = json.load(f)
        from rectified_flow_vision_b200 import _build
        per_img = d["dram_bytes_per_image"].get(kind)
        if per_img is None:
            return None, f"{os.path.basename(NCU_TRAFFIC_FILE)} has no entry for {kind}"
        note = (f"dram__bytes_read.sum + dram__bytes_write.sum of the {kind} launches of one forward, ncu --set full capture "
                f"profiles/{os.path.basename(NCU_TRAFFIC_FILE)} (micro-batch {d['micro_batch']}), scaled to micro-batch {micro_batch}")
        if d.get("library_digest") != _build._digest():
            note += "; captured with an EARLIER build of the library (kernel sources changed since)"
        return per_img * micro_batch, note
    except Exception as ex:  # noqa: BLE001
        return None, f"no ncu traffic capture available ({type(ex).__name__})"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm (oracle port on the host cores)
# ------------------------------------------------------------------------------------------------------------
def cpu_pairs_per_sec(sample_images: int, sample_steps: int, repeats: int = 1):
    """Time the reference's CPU implementation of the path on the host cores, all threads, on a bounded sample:
    `sample_images` noises integrated for `sample_steps` Euler steps (cost is linear in steps -- reference CSV,
    results/benchmark_results.csv:2-9 -- so pairs/s at 100 steps = images*steps/s / 100).

    kind "reference": the UNMODIFIED reference package vendored in oracle/_ref (oracle/build_ref.py), driven through its own
    public API `BaseFlowModel.sample(noise=..., num_steps=...)` (models/base_flow.py:133-177); kind "port": the functional
    restatement oracle/torch_port.py, only when oracle/_ref was never built."""
    import torch
    from oracle import build_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    noise = torch.randn(sample_images, CH, IMAGE, IMAGE, generator=torch.Generator().manual_seed(42))
    if build_ref.available():
        ref = build_ref.import_ref()
        torch.manual_seed(0)
        m = ref.BaseFlowModel(image_size=IMAGE, device="cpu")
        m.eval()
        run, kind = (lambda x, n: m.sample(noise=x, num_steps=n)), "reference"
        what = "unmodified reference (oracle/_ref, models/base_flow.py BaseFlowModel.sample) on the host CPU, fp32"
    else:
        from oracle import torch_port
        import rectified_flow_vision_b200 as pkg
        torch.manual_seed(0)
        m = pkg.BaseFlowModel(device="cpu")
        P = {k: v.detach() for k, v in m.state_dict().items()}
        run, kind = (lambda x, n: torch_port.euler_sample(P, x, n)), "port"
        what = "functional-PyTorch fp32 port of the reference path (oracle/torch_port.py) on the host CPU"
    run(noise[:2], 1)  # warm-up
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        run(noise, sample_steps)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    img_steps_per_s = sample_images * sample_steps / best
    return img_steps_per_s / EULER_STEPS, cores, best, kind, what


def torch_eager_gpu_image_steps_per_sec(dev, batch: int = 256, steps: int = 2):
    """Same-box GPU baseline: the functional-PyTorch port of the reference path run by PyTorch eager / cuDNN on this
    GPU -- fp32 with TF32 off, fp32 with TF32 on, and bf16 autocast.  A baseline only: nothing of it is on the product path."""
    import torch
    from oracle import torch_port
    import rectified_flow_vision_b200 as pkg
    torch.manual_seed(0)
    m = pkg.BaseFlowModel(device="cpu")
    P = {k: v.detach().to(dev) for k, v in m.state_dict().items()}
    noise = torch.randn(batch, CH, IMAGE, IMAGE, generator=torch.Generator().manual_seed(7)).to(dev)
    out = {}

    def run():
        x = noise
        dt = 1.0 / steps
        for i in range(steps):
            t = torch.full((batch,), i * dt, device=dev)
            x = x + torch_port.unet_forward(P, x, t) * dt
        return x

    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        with torch.device(dev):
            for name, tf32, amp in (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    run()
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(2):
                        run()
                    b.record()
                    torch.cuda.synchronize()
                out[name] = batch * steps * 2 / (a.elapsed_time(b) / 1e3)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out["unit"] = "image-steps/s"
    out["sample"] = f"batch {batch} x {steps} Euler steps, oracle/torch_port.py on cuda via PyTorch eager + cuDNN"
    return out


def torch_eager_gpu_train_images_per_sec(dev, batch: int = 256):
    """Same-box GPU baseline of the training step: the reference's step body (interpolation, forward, mse, backward,
    clip_grad_norm_(1.0), torch.optim.AdamW) on the functional port, PyTorch eager + cuDNN on this GPU, with TF32 and with
    bf16 autocast.  Baseline only: nothing of it is on the product path."""
    import torch
    from oracle import torch_port
    import rectified_flow_vision_b200 as pkg
    torch.manual_seed(0)
    m = pkg.BaseFlowModel(device="cpu")
    g = torch.Generator().manual_seed(3)
    x0, x1, t = (torch.randn(batch, CH, IMAGE, IMAGE, generator=g).to(dev), torch.randn(batch, CH, IMAGE, IMAGE, generator=g).to(dev),
                 torch.rand(batch, generator=g).to(dev))
    out = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        with torch.device(dev):
            for name, amp in (("tf32", False), ("bf16_autocast", True)):
                torch.backends.cuda.matmul.allow_tf32 = True
                torch.backends.cudnn.allow_tf32 = True
                P = {k: v.detach().clone().to(dev).requires_grad_(True) for k, v in m.state_dict().items()}
                opt = torch.optim.AdamW(list(P.values()), lr=1e-4)
                times = []
                for step in range(4):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    tt = t.view(-1, 1, 1, 1)
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                        pred = torch_port.unet_forward_grad(P, (1 - tt) * x0 + tt * x1, t)
                        loss = torch.nn.functional.mse_loss(pred.float(), x1 - x0)
                    opt.zero_grad()
                    loss.backward()
                    torch.nn.utils.clip_grad_norm_(list(P.values()), 1.0)
                    opt.step()
                    torch.cuda.synchronize()
                    times.append(time.perf_counter() - t0)
                out[name] = batch / min(times[1:])
                del P, opt
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out["unit"] = "images/s"
    out["sample"] = f"one optimizer step at batch {batch}: autograd over oracle/torch_port.py + clip_grad_norm_ + torch.optim.AdamW, PyTorch eager on cuda"
    return out


def cpu_train_images_per_sec(batch: int):
    import torch
    from oracle import train_oracle
    import rectified_flow_vision_b200 as pkg
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    m = pkg.BaseFlowModel(device="cpu")
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x0, x1, t = torch.randn(batch, CH, IMAGE, IMAGE, generator=g), torch.randn(batch, CH, IMAGE, IMAGE, generator=g), torch.rand(batch, generator=g)
    state = {}
    best = None
    for step in range(3):  # first step is the warm-up
        t0 = time.perf_counter()
        _, grads = train_oracle.loss_and_grads(P, x0, x1, t)
        train_oracle.adamw_step(P, grads, state, step + 1)
        dt = time.perf_counter() - t0
        if step > 0:
            best = dt if best is None else min(best, dt)
    return {"value": batch / best, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"one optimizer step at batch {batch} ({best:.2f} s): oracle/train_oracle.py (autograd over the fp32 port + restated clip/AdamW)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    imgs, steps = 16, 4
    times = []
    for i in range(args.warmup + args.steps):
        _, cores, dt, kind, what = cpu_pairs_per_sec(imgs, steps)
        if i >= args.warmup:
            times.append(dt)
    mean_dt = sum(times) / len(times)
    value = (imgs * steps / mean_dt) / EULER_STEPS
    sample = f"{imgs} seeded noises x {steps} Euler steps per step, scaled linearly to {EULER_STEPS} steps; {what}"
    line = {"impl": "reference", "metric": "reflow_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, default_mb()),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def default_mb():
    from rectified_flow_vision_b200.engine import default_micro_batch
    return default_micro_batch(IMAGE)


def workload_config(args, micro_batch):
    return {"workload": f"reflow pair generation: default UNet 64x64x3 (base_flow_final.pt architecture, seeded random-init "
                        f"weights), {EULER_STEPS}-step Euler, {args.pairs_per_step} seeded noises per GPU per step out of the "
                        f"{TOTAL_PAIRS}-pair job, batch-sharded over {args.gpus} GPU(s)",
            "pairs_per_step_per_gpu": args.pairs_per_step, "euler_steps": EULER_STEPS, "image": [CH, IMAGE, IMAGE],
            "micro_batch": micro_batch, "parallelism": f"dp{args.gpus} (batch shards, no data-path collective)",
            "l2_policy": "inputs larger than L2: every Euler step streams > 2 GB of activations per micro-batch (L2 = 126 MB); the "
                         "noise comes from 4 rotating 100 MB buffers"}


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import rectified_flow_vision_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    torch.manual_seed(0)  # identical replicas on every rank
    model = pkg.BaseFlowModel(device=f"cuda:{local}")
    model.eval()
    eng = model._engine()
    P = args.pairs_per_step
    gen = torch.Generator().manual_seed(42 + rank)
    n_bufs = args.warmup + args.steps
    host_noise = [torch.randn(P, CH, IMAGE, IMAGE, generator=gen).pin_memory() for _ in range(min(n_bufs, 4))]
    dev_noise = [h.to(dev) for h in host_noise]

    # ---- value: inputs resident in HBM, device-timed ----
    for i in range(args.warmup):
        eng.euler_sample(dev_noise[i % len(dev_noise)], EULER_STEPS)
    barrier()
    eng.launch_count(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        eng.euler_sample(dev_noise[(args.warmup + i) % len(dev_noise)], EULER_STEPS)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count(reset=True)
    value = world * P * args.steps / (ms / 1e3)

    # ---- e2e: the public API with host buffers (H2D + D2H inside the timed region) ----
    e2e_steps = max(1, min(args.steps, 6))
    out = pkg.generate_reflow_pairs(model, num_pairs=P, num_steps=EULER_STEPS, noise=host_noise[0])  # warm
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        x0, x1 = pkg.generate_reflow_pairs(model, num_pairs=P, num_steps=EULER_STEPS, noise=host_noise[i % len(host_noise)])
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * P * e2e_steps / e2e_s
    img_bytes = P * CH * IMAGE * IMAGE * 4
    del out, x0, x1

    # ---- e2e with the job's final gather (N > 1): every rank integrates its shard of ONE host-seeded noise tensor and the
    #      results are all-gathered over NCCL into row order on every rank's host (dist.generate_reflow_pairs_sharded) ----
    e2e_gather = None
    if world > 1:
        from rectified_flow_vision_b200 import dist as rdist
        job_noise = rdist.seeded_noise(world * P, CH, IMAGE, seed=4242).pin_memory()   # identical on every rank
        # the gathered result lands in ONE pinned host buffer the job re-uses (a fresh 805 MB pinned tensor per call costs ~0.3 s
        # of page-locking at 8 ranks -- the caller of a repeated gather passes `out=`, as here)
        g_out = torch.empty((world * P, CH, IMAGE, IMAGE), dtype=torch.float32, pin_memory=True)
        rdist.generate_reflow_pairs_sharded(model, world * P, EULER_STEPS, noise=job_noise, gather=True, out=g_out)   # warm (NCCL channels)
        barrier()
        t0 = time.perf_counter()
        g_steps = max(1, min(args.steps, 2))
        for _ in range(g_steps):
            gx0, gx1 = rdist.generate_reflow_pairs_sharded(model, world * P, EULER_STEPS, noise=job_noise, gather=True, out=g_out)
        torch.cuda.synchronize()
        g_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        assert gx1.shape[0] == world * P
        e2e_gather = {"value": world * P * g_steps / g_s, "unit": "pairs/s", "steps": g_steps,
                      "all_gather_bytes_per_rank_per_step": int(world * img_bytes),
                      "api": "dist.generate_reflow_pairs_sharded(model, num_pairs, 100, noise=<host tensor>, gather=True): shard -> "
                             "integrate -> ONE all_gather_into_tensor over NCCL -> full (x0, x1) on every rank's host"}
        del gx0, gx1, job_noise, g_out

    # ---- BASELINE.json configs[4]: the config.yaml UNet at 128x128, seeded random-init weights, 8-step Euler, 128 images per
    #      GPU (batch 1,024 over 8 GPUs), inputs resident in HBM ----
    cfg5 = None
    if not args.no_128:
        torch.manual_seed(0)
        m128 = pkg.BaseFlowModel(image_size=128, device=f"cuda:{local}")
        m128.eval()
        e128 = m128._engine(128)
        nz = torch.randn(128, CH, 128, 128, generator=torch.Generator().manual_seed(7 + rank)).to(dev)
        for _ in range(2):
            e128.euler_sample(nz, 8)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record()
        for _ in range(reps):
            e128.euler_sample(nz, 8)
        b.record()
        barrier()
        ms128 = max_over_ranks(a.elapsed_time(b)) / reps
        fl128 = e128.flops_per_image()
        pk_ = peaks()
        tf = 128 * 8 * fl128 / (ms128 / 1e3) / 1e12
        cfg5 = {"images_per_sec": world * 128 / (ms128 / 1e3), "ms_per_batch": ms128, "batch_per_gpu": 128, "euler_steps": 8,
                "image": [CH, 128, 128], "gflop_per_image_step": fl128 / 1e9, "tflops_per_gpu": tf,
                "frac_of_burst": tf / pk_["burst"], "frac_of_sustained": tf / pk_["sustained"],
                "what": "config.yaml UNet (64 ch, mult [1,2,4], 2 res blocks) at 128x128, seeded random-init weights, 8-step Euler"}
        del e128, m128, nz
        torch.cuda.empty_cache()

    # ---- reflow training step (BASELINE.json configs[3]): fwd + bwd + clip + AdamW on synthetic pairs, data parallel:
    #      per-GPU batch fixed (weak scaling), ONE all-reduce of the flat fp32 gradient buffer per step over NCCL ----
    train = None
    if not args.no_train:
        from rectified_flow_vision_b200.training import NativeTrainer
        tb = args.train_batch
        tmodel = pkg.RectifiedFlowModel(device=f"cuda:{local}")   # same seed on every rank: identical replicas
        tmodel.train()
        tr = NativeTrainer(tmodel, lr=1e-4, micro_batch=min(tb, 256))
        tg = torch.Generator().manual_seed(1000 + rank)
        tx0 = torch.randn(tb, CH, IMAGE, IMAGE, generator=tg).to(dev)
        tx1 = torch.randn(tb, CH, IMAGE, IMAGE, generator=tg).to(dev)
        tt = torch.rand(tb, generator=tg).to(dev)
        for _ in range(3):
            tr.step(tx0, tx1, tt)
        barrier()
        teng = tmodel.velocity_net.train_engine(IMAGE, dev)
        teng.launch_count(reset=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tsteps = 5
        a.record()
        for _ in range(tsteps):
            loss = tr.step(tx0, tx1, tt)
        b.record()
        barrier()
        tms = max_over_ranks(a.elapsed_time(b)) / tsteps
        train = {"images_per_sec": world * tb / (tms / 1e3), "ms_per_step": tms, "batch_per_gpu": tb, "dropout": 0.1,
                 "loss": float(loss.item()), "grad_allreduce_bytes": int(teng.grad_buffer().numel() * 4) if world > 1 else 0,
                 "gpu_launches_per_step": int(teng.launch_count(reset=True)) // tsteps,
                 "tflops_at_3x_forward": world * tb * 3 * FLOPS_PER_IMG_STEP / (tms / 1e3) / 1e12,
                 "what": "train_rectified_flow step body: x_t interpolation, UNet fwd, MSE, bwd, clip_grad_norm_(1.0), AdamW; "
                         "synthetic N(0,1) pairs, t ~ U[0,1), seeded; gradients all-reduced (SUM) then scaled 1/world"}
        if rank == 0:
            # per-kernel-class split of one training step and the rooflines of its dominant tensor / HBM kernels, measured on a
            # second engine that runs the backward pass on ONE stream (RFV_FLAG_ONE_STREAM): with the production two-stream
            # schedule the weight-gradient and GroupNorm-backward kernels share the GPU and per-kernel event times overlap
            from rectified_flow_vision_b200 import engine as _E
            peng = _E.Engine(tmodel.velocity_net.arch(), IMAGE, dev, micro_batch=min(tb, 256), train=True, flags=2048)
            peng.sync_weights(tmodel.velocity_net)
            peng.zero_grad()
            peng.train_accumulate(tx0, tx1, tt, dropout_p=0.1, seed=1)
            peng.set_profiling(True)
            peng.zero_grad()
            peng.train_accumulate(tx0, tx1, tt, dropout_p=0.1, seed=12345)
            rep = peng.profile_report()
            peng.set_profiling(False)
            del peng
            ksplit = {}
            for ln in rep.strip().splitlines():
                key, ms_tot, n, fl_img, by_img = ln.split("\t")
                kind = key.split(" ", 1)[0] + ("/bwd" if " bwd:" in key else "")
                k = ksplit.setdefault(kind, {"ms": 0.0, "launches": 0, "gflop_per_image": 0.0, "mbyte_per_image": 0.0})
                k["mbyte_per_image"] += float(by_img) / 1e6
                k["ms"] += float(ms_tot)
                k["launches"] += int(n)
                k["gflop_per_image"] += float(fl_img) / 1e9
            train["kernel_split_ms_per_step"] = ksplit
            pk_ = peaks()
            if "wgrad_umma/bwd" in ksplit:
                w = ksplit["wgrad_umma/bwd"]
                ach = w["gflop_per_image"] * 1e9 * tb / (w["ms"] / 1e3) / 1e12
                train["roofline"] = {"bound": "tensor", "kernel": "wgrad_umma_kernel (tcgen05 weight gradients, MN-major operands; all launches of one step)",
                                     "achieved": ach, "peak": pk_["sustained"], "unit": "TFLOP/s", "frac": ach / pk_["sustained"],
                                     "frac_of_burst": ach / pk_["burst"], "peak_source": pk_["source"], "traffic": None,
                                     "flops_per_step": w["gflop_per_image"] * 1e9 * tb, "ms_per_step": w["ms"]}
            if "gn_bwd/bwd" in ksplit:
                # algorithmic bytes of the launches that ran (engine profile report): single-pass sites 6 B per element (x, dy in;
                # dx out), two-pass sites 4 + 6 B; addends extra
                gb = ksplit["gn_bwd/bwd"]["mbyte_per_image"] * 1e6 * tb
                ach = gb / (ksplit["gn_bwd/bwd"]["ms"] / 1e3) / 1e9
                train["roofline_hbm"] = {"bound": "hbm", "kernel": "gn_bwd_fused_kernel / gn_bwd_kernel<false/true> (GroupNorm+SiLU+dropout backward, %d launches of one step)" % ksplit["gn_bwd/bwd"]["launches"],
                                         "achieved": ach, "peak": pk_["hbm"], "unit": "GB/s", "frac": ach / pk_["hbm"],
                                         "bytes_per_step": gb, "ms_per_step": ksplit["gn_bwd/bwd"]["ms"]}
        del tr, tmodel, teng
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    line = {"metric": "reflow_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, default_mb()),
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": img_bytes, "d2h_bytes_per_step": img_bytes,
                    "api": "generate_reflow_pairs(model, num_pairs, num_steps=100, noise=<pinned host tensor>)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "unet_flops_fraction_of_peak": {
                "achieved_tflops": value * EULER_STEPS * FLOPS_PER_IMG_STEP / 1e12 / world,
                "of_sustained": value * EULER_STEPS * FLOPS_PER_IMG_STEP / 1e12 / world / pk["sustained"],
                "of_burst": value * EULER_STEPS * FLOPS_PER_IMG_STEP / 1e12 / world / pk["burst"], "peaks": pk["source"]}}
    if e2e_gather is not None:
        # N > 1: the headline e2e is the variant that ends with the job's final gather; the per-rank (no collective) number stays
        line["e2e_no_gather"] = dict(line["e2e"])
        line["e2e"] = {**e2e_gather, "h2d_bytes_per_step": img_bytes, "d2h_bytes_per_step": int(world * img_bytes)}
    if cfg5 is not None:
        line["config5_128x128"] = cfg5
    summary = {"pairs_per_sec": round(value, 1), "e2e_pairs_per_sec": round(line["e2e"]["value"], 1)}
    if e2e_gather is not None:
        summary["e2e_no_gather_pairs_per_sec"] = round(e2e_value, 1)
    if train is not None:
        summary.update(train_images_per_sec=round(train["images_per_sec"], 1), train_ms_per_step=round(train["ms_per_step"], 3),
                       train_allreduce_bytes=train["grad_allreduce_bytes"])
    if cfg5 is not None:
        summary.update(cfg5_128px_images_per_sec=round(cfg5["images_per_sec"], 1), cfg5_frac_of_burst=round(cfg5["frac_of_burst"], 3))
    line["summary"] = summary
    if train is not None:
        line["train_step"] = train

    if world == 1:
        # ---- per-kernel-class split + roofline of the dominant kernel, measured live with CUDA events ----
        mb = eng.micro_batch
        x = dev_noise[0][:mb].clone()
        eng.set_profiling(True)
        for _ in range(3):
            eng.euler_sample(x, 1)
        rep = eng.profile_report()
        eng.set_profiling(False)
        kinds = {}
        for ln in rep.strip().splitlines():
            key, ms_tot, n, fl_img, by_img = ln.split("\t")
            kind, label = key.split(" ", 1)
            k = kinds.setdefault(kind, {"ms": 0.0, "launches": 0, "gflop_per_image": 0.0, "mbyte_per_image": 0.0})
            k["mbyte_per_image"] += float(by_img) / 1e6
            k["ms"] += float(ms_tot) / 3.0
            k["launches"] += int(n) // 3
            k["gflop_per_image"] += float(fl_img) / 1e9
        tot_ms = sum(k["ms"] for k in kinds.values())
        for k in kinds.values():
            k["share"] = k["ms"] / tot_ms
        line["kernel_split_ms_per_forward"] = {"micro_batch": mb, **{k: v for k, v in kinds.items()}}
        tc = {k: kinds[k] for k in ("conv_wa", "conv_halo", "conv_umma") if k in kinds}
        if tc:
            # dominant kernel = the tcgen05 conv class with the most time in a forward; algorithmic FLOPs (2*MACs,
            # rfv_profile_report) of the convolutions it executed / CUDA-event time of its launches
            names = {"conv_wa": "conv_wa_kernel<PAIR,FUSE> (tcgen05 3x3 convs, weights as the A operand, up to 256 pixels as B; all launches of one forward)",
                     "conv_halo": "conv_halo_kernel<BN> (tcgen05 3x3 convs with halo reuse, all launches of one forward)",
                     "conv_umma": "conv_umma_kernel<BN> (tcgen05 implicit-GEMM convs, all launches of one forward)"}
            dom = max(tc, key=lambda k: tc[k]["ms"])
            fl = tc[dom]["gflop_per_image"] * 1e9 * mb
            ach = fl / (tc[dom]["ms"] / 1e3) / 1e12
            fl_all = sum(v["gflop_per_image"] for v in tc.values()) * 1e9 * mb
            ach_all = fl_all / (sum(v["ms"] for v in tc.values()) / 1e3) / 1e12
            # the split is measured in a 3-forward burst with per-launch events (kernels timed in isolation): burst peak
            traffic, traffic_note = ncu_traffic(dom, mb)
            line["roofline"] = {"bound": "tensor", "kernel": names[dom],
                                "achieved": ach, "peak": pk["burst"], "unit": "TFLOP/s", "frac": ach / pk["burst"],
                                "frac_of_sustained": ach / pk["sustained"],
                                "peak_source": pk["source"] + " (burst: the split is event-timed per launch in a 3-forward burst)",
                                "traffic": traffic, "traffic_note": traffic_note,
                                "flops_per_forward": fl, "ms_per_forward": tc[dom]["ms"],
                                "all_tcgen05_convs": {"achieved": ach_all, "frac": ach_all / pk["burst"],
                                                      "frac_of_sustained": ach_all / pk["sustained"]}}
        if "gn_apply" in kinds:
            # second-largest kernel class, HBM-bound: algorithmic bytes = elements x (2 B read + 2 B written)
            # (bytes of the launches that ran, from the engine's profile report: sites fused into their conv launch nothing)
            gb = kinds["gn_apply"]["mbyte_per_image"] * 1e6 * mb
            ach = gb / (kinds["gn_apply"]["ms"] / 1e3) / 1e9
            line["roofline_hbm"] = {"bound": "hbm", "kernel": "gn_apply_kernel (GroupNorm + SiLU + virtual concat, %d launches of one forward)" % kinds["gn_apply"]["launches"],
                                    "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                                    "peak_source": pk["source"], "bytes_per_forward": gb, "ms_per_forward": kinds["gn_apply"]["ms"]}
        # ---- sampling throughput, configs[0] (B=64) and configs[2] (B=4096) ----
        samp = {}
        for bsz, reps in ((64, 5), (4096, 2)):
            nz = torch.randn(bsz, CH, IMAGE, IMAGE, generator=torch.Generator().manual_seed(1)).to(dev)
            for steps in (1, 2, 4, 8):
                if bsz == 4096 and steps > 2:
                    continue
                eng.euler_sample(nz, steps)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    eng.euler_sample(nz, steps)
                b.record()
                torch.cuda.synchronize()
                samp[f"b{bsz}_steps{steps}"] = bsz * reps / (a.elapsed_time(b) / 1e3)
        line["sampling_images_per_sec"] = samp
        # ---- CPU baseline: the oracle port on this host's cores, bounded sample ----
        try:
            v, cores, secs, kind, what = cpu_pairs_per_sec(16, 4)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind,
                                    "sample": f"16 seeded noises x 4 Euler steps ({secs:.1f} s), scaled linearly to {EULER_STEPS} steps; {what}"}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "error": str(ex)}
        try:
            line["torch_eager_gpu_baseline"] = torch_eager_gpu_image_steps_per_sec(dev)
            line["torch_eager_gpu_baseline"]["ours_image_steps_per_sec"] = value * EULER_STEPS
        except Exception as ex:  # noqa: BLE001
            line["torch_eager_gpu_baseline"] = {"error": str(ex)[:200]}
        if train is not None:
            try:   # the training step of the oracle (autograd over the functional port + restated AdamW) on the host cores
                train["cpu_baseline"] = cpu_train_images_per_sec(8)
            except Exception as ex:  # noqa: BLE001
                train["cpu_baseline"] = {"value": None, "error": str(ex)}
            try:
                train["torch_eager_gpu_baseline"] = torch_eager_gpu_train_images_per_sec(dev)
            except Exception as ex:  # noqa: BLE001
                train["torch_eager_gpu_baseline"] = {"error": str(ex)[:200]}
    line["summary_tail"] = line["summary"]   # the same short summary again as the LAST key: log tails keep the end of the line
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=2048)
    ap.add_argument("--train-batch", type=int, default=256, help="per-GPU batch of the training-step measurement")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement")
    ap.add_argument("--no-128", action="store_true", help="skip the 128x128 (configs[4]) measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
