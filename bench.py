#!/usr/bin/env python
"""Benchmark of the SemiVL training hot path on B200 (BASELINE.json metric: training images/sec, 512x512, ViT-B/16).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU PyTorch path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): VOC 21-class synthetic, 512x512, ViT-B/16 + VLG head, per-GPU batch 16, one supervised
training step = encoder fwd -> head fwd -> fused upsample+CE -> head bwd -> encoder bwd -> (NCCL grad all-reduce) -> AdamW.
`--workload semivl` times the full SemiVL consistency step (teacher + MaskCLIP + 5-way student head) instead.
Prints ONE JSON line (rank 0).
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

CROP, NCLASS, BATCH = 512, 21, 16
# algorithmic GFLOP per image of the supervised step at 512^2 / N=21 (FlopCounterMode on the reference: matmul+conv, 2*MAC)
GF_PER_IMG_SUPERVISED = 831.5
GF_PER_UNIT_SEMIVL = 4542.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="supervised", choices=["supervised", "semivl"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--crop", type=int, default=CROP)
    ap.add_argument("--nclass", type=int, default=NCLASS)
    ap.add_argument("--precise", action="store_true", help="split-bf16 parity mode instead of the bf16 throughput mode")
    ap.add_argument("--graph-multi", action="store_true", help="N > 1: capture the step (NCCL exchange included) in a CUDA graph as well")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the ~690 kernels of the step eagerly instead of replaying the captured CUDA graph (N=1)")
    ap.add_argument("--cpu-crop", type=int, default=CROP)
    return ap.parse_args()


def model_cfg(crop, nclass, precise):
    return dict(model='mmseg.vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb', nclass=nclass, crop_size=crop, dataset='pascal' if nclass == 21 else 'ade',
                text_embedding_variant='single', mcc_text='single', pl_text='single', clip_encoder='mcvit16', disable_dropout=True,
                fp_rate=0.5, model_args=dict(pretrained=None), clip_encoder_args=dict(pretrained=None), precise=precise)


def synth_batch(torch, b, crop, nclass, seed, device, semivl):
    """Synthetic inputs of SURVEY.md §8d: randn images, labels with a 5% ignore region, pad-strip ignore masks, CutMix boxes."""
    g = torch.Generator().manual_seed(seed)
    r = lambda: torch.randn(b, 3, crop, crop, generator=g)
    mask = torch.randint(0, nclass, (b, crop, crop), generator=g)
    mask[:, : crop // 5, : crop // 4] = 255
    out = dict(img_x=r(), mask_x=mask)
    if semivl:
        for k in ("img_w", "img_s1", "img_s2", "img_w_other", "img_s1_other", "img_s2_other"):
            out[k] = r()
        ign = torch.zeros(b, crop, crop, dtype=torch.long)
        ign[:, -crop // 10:, :] = 255
        ign_o = torch.zeros(b, crop, crop, dtype=torch.long)
        ign_o[:, :, -crop // 12:] = 255
        def box(frac):
            m = torch.zeros(b, crop, crop)
            h = int(crop * frac)
            m[::2, crop // 5: crop // 5 + h, crop // 4: crop // 4 + h] = 1
            return m
        out.update(ignore_mask=ign, ignore_mask_other=ign_o, mix1=box(0.5), mix2=box(0.3))
    return {k: (v.pin_memory() if device != "cpu" else v) for k, v in out.items()}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max((float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()), default=None), "reasons": reasons}


def cpu_reference_rate(args, steps, warmup, crop, b=1):
    """The reference's CPU PyTorch path (oracle port, fp32) on the host cores: supervised step fwd + CE + bwd at batch b."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from oracle import semivl_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mc = O.ModelCfg(img_size=crop, num_classes=args.nclass)
    sd = O.fixture_state_dict(O.param_shapes(mc, with_clip_encoder=False), seed=0)
    p = {k: v.clone().requires_grad_(("attn" in k or "pos_embed" in k) if k.startswith("backbone.") else True) for k, v in sd.items()}
    tname = "voc12_wbg_single" if args.nclass == 21 else "ade_single"
    text = torch.from_numpy(np.load(os.path.join(ROOT, "semivl_b200", "configs", "_base_", "datasets", "text_embedding", tname + ".npy")))
    batch = synth_batch(torch, b, crop, args.nclass, 1234, "cpu", False)
    # the reference's optimizer (experiments.py:246-255 via mmcv's constructor): torch.optim.AdamW, backbone lr x0.01, head lr x10
    train = {k: v for k, v in p.items() if v.requires_grad}
    opt = torch.optim.AdamW([dict(params=[v for k, v in train.items() if k.startswith("backbone.")], lr=1e-4 * 0.01),
                             dict(params=[v for k, v in train.items() if not k.startswith("backbone.")], lr=1e-4 * 10.0)], lr=1e-4, weight_decay=0.01)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = O.supervised_step_loss(batch["img_x"], batch["mask_x"], p, text, mc)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return b / (ms / 1e3), ms, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 16)), max(1, min(args.warmup, 2))        # ~0.7 s per CPU step at 512^2: bounded to ~12 s
    value, ms, cores = cpu_reference_rate(args, steps, warmup, args.cpu_crop)
    sample = f"supervised step fwd+CE+bwd+AdamW at batch 1, {args.cpu_crop}x{args.cpu_crop}, N={args.nclass}, {steps} timed steps after {warmup} warm-up"
    line = {"impl": "reference", "metric": "training images/sec", "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": f"VOC {args.nclass}-class synthetic {args.cpu_crop}x{args.cpu_crop} ViT-B/16+VLG head supervised step",
                                            "batch_per_step": 1, "note": "reference CPU path = oracle port of the reference (pure PyTorch fp32)"},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective: park fd 1 on stderr until the result line is due, so
        # that rank 0's stdout carries exactly one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    from semivl_b200 import lib as L
    from semivl_b200 import ops
    from semivl_b200.model import build_model
    from semivl_b200.train import OptimCfg, Trainer
    L.check_device()
    semivl = args.workload == "semivl"
    torch.manual_seed(0)
    model = build_model(model_cfg(args.crop, args.nclass, args.precise)).to(dev)
    tr = Trainer(model, OptimCfg(lr=1e-4, total_iters=100000))
    b = args.batch
    host = synth_batch(torch, b, args.crop, args.nclass, 1234 + rank, "cuda", semivl)
    resident = {k: v.to(dev) for k, v in host.items()}

    use_graph = not semivl and not args.no_graph and args.crop % 16 == 0 and (world == 1 or args.graph_multi)

    def step(batch):
        if semivl:
            return tr.semivl_step(batch)[0]
        if use_graph and ops.PROFILE is None:         # the instrumented steps (events around each contraction) run eagerly
            return tr.graphed_supervised_step(batch["img_x"], batch["mask_x"])
        return tr.supervised_step(batch["img_x"], batch["mask_x"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 12)):       # W untimed steps, and never fewer than 12: graph capture + clock ramp of a cold box
        step(resident)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- device-timed region: inputs resident in HBM
    l0 = L.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(resident)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (L.launches - l0) // args.steps
    # ---- end-to-end region: host (pinned) buffers; every step's inputs cross PCIe inside the timed region and the loss is read
    # back every step.  The copies are double-buffered on a copy stream (what a pinned-memory data loader does): the H2D of step
    # i+1 runs under the kernels of step i, each step waits for its own inputs.
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    stages = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])              # the step that last read this staging buffer has finished
            for k, v in host.items():
                stages[i % 2][k].copy_(v, non_blocking=True)
            ready[i % 2].record(copy_stream)

    barrier()
    for ev in consumed:
        ev.record()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    upload(0)
    # every step's loss crosses PCIe into pinned memory; the host reads it one step behind (what an asynchronous logger does), so the
    # launch of step i+1 is not held back by the read-back of step i
    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    lv = float("nan")
    for i in range(args.steps):
        if i + 1 < args.steps:
            upload(i + 1)
        torch.cuda.current_stream().wait_event(ready[i % 2])
        out = step(stages[i % 2])
        consumed[i % 2].record()
        loss_host[i % 2].copy_(out.reshape(1), non_blocking=True)
        loss_ready[i % 2].record()
        if i > 0:
            loss_ready[(i - 1) % 2].synchronize()
            lv = float(loss_host[(i - 1) % 2][0])
    loss_ready[(args.steps - 1) % 2].synchronize()
    lv = float(loss_host[(args.steps - 1) % 2][0])
    f1.record()
    barrier()
    ms_e2e = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3) / args.steps
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    # ---- instrumented steps: CUDA events (on the launching stream) around every tensor-core contraction launch
    ops.PROFILE = []
    n_inst = 2
    for _ in range(n_inst):
        step(resident)
    torch.cuda.synchronize()
    prof = [(a.elapsed_time(bv), f, lab) for a, bv, f, lab in ops.PROFILE]
    ops.PROFILE = None
    gemm_ms = sum(t_ for t_, _, _ in prof) / n_inst
    gemm_flops = sum(f for _, f, _ in prof) / n_inst
    n_gemm = len(prof) // n_inst
    by_shape = {}
    for t_, f, lab in prof:
        e = by_shape.setdefault(lab, [0, 0.0, 0.0])
        e[0] += 1
        e[1] += t_
        e[2] += f
    dom_label, dom = max(by_shape.items(), key=lambda kv: kv[1][1])         # the launch shape with the largest share of the step
    dom_n, dom_ms, dom_fl = dom[0] / n_inst, dom[1] / n_inst, dom[2] / n_inst
    if os.environ.get("SVL_PROFILE_DUMP") and rank == 0:
        with open(os.environ["SVL_PROFILE_DUMP"], "w") as fh:
            for t_ms, f, lab in prof[:n_gemm]:
                fh.write(f"{t_ms * 1e3:10.1f} us {f / t_ms / 1e9 if t_ms > 0 else 0:8.1f} TF/s  {lab}\n")
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank != 0:
        teardown(world, tr)
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    dom_tf = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    traffic = {}
    try:        # DRAM bytes per launch of the dominant shape, from the committed `ncu --set full` capture (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json"))).get(dom_label, {})
    except Exception:
        pass
    unit_gf = GF_PER_UNIT_SEMIVL if semivl else GF_PER_IMG_SUPERVISED
    step_tf = unit_gf * b / 1e3
    value = world * b / (ms / 1e3)
    line = {
        "metric": "training images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3" if args.precise else "bf16",
        "data": "synthetic",
        "config": {"workload": (f"VOC {args.nclass}-class synthetic {args.crop}x{args.crop} ViT-B/16+VLG head, "
                                + ("full SemiVL consistency step" if semivl else "supervised step fwd+bwd+AdamW")),
                   "batch_per_gpu": b, "global_batch": b * world, "parallelism": f"dp{world}",
                   "launch": "one CUDA-graph replay per step" if use_graph else "eager kernel launches", "l2": "per-step working set (GBs of activations) >> 126 MB L2",
                   "weights": "random init (reference init_weights)", "loss": float(loss.item())},
        "e2e": {"value": world * b / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e, "last_loss": lv,
                "how": "pinned host inputs uploaded every step on a copy stream (double-buffered), every step's loss copied to pinned host memory and read one step behind"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary() if sampler else None,
        "roofline": {"bound": "tensor", "kernel": f"gemm_kernel [{dom_label}] (persistent TMA + tcgen05 GEMM, csrc/gemm.cu)",
                     "achieved": dom_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": dom_tf / peak_tf if peak_tf else None,
                     "traffic": traffic.get("dram_bytes_per_launch"), "traffic_source": traffic.get("source"),
                     "algorithmic_flops_per_launch": dom_fl / dom_n, "algorithmic_bytes_per_launch": traffic.get("algorithmic_bytes_per_launch"),
                     "avg_launch_us": 1e3 * dom_ms / dom_n, "launches_per_step": dom_n, "share_of_step": dom_ms / ms if ms else None,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                     "measured_on": f"{n_inst} instrumented steps after the timed region (CUDA events on the launching stream around every contraction launch)",
                     "engine": {"what": "all gemm_kernel / wgrad_kernel / wgrad_strip_kernel launches of the step", "launches": n_gemm,
                                "achieved": achieved_tf, "frac": achieved_tf / peak_tf if peak_tf else None, "ms_per_step": gemm_ms,
                                "share_of_step": gemm_ms / ms if ms else None},
                     "whole_step_algorithmic_tflops": step_tf / (ms / 1e3), "whole_step_frac": step_tf / (ms / 1e3) / peak_tf},
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            v, cms, cores = cpu_reference_rate(args, 14, 2, args.cpu_crop)          # ~10 s of host work
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"oracle port of the reference (CPU PyTorch fp32), supervised step fwd+CE+bwd+AdamW at batch 1, "
                                              f"{args.cpu_crop}x{args.cpu_crop}, N={args.nclass}: 14 timed steps after 2 warm-up ({cms:.0f} ms/step)"}
        except Exception as e:      # the baseline must never sink the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    teardown(world, tr)


def teardown(world, tr):
    """Release the captured graph before the communicator; a watchdog ends the process if the NCCL teardown stalls (seen once with a
    live captured graph at N=2) -- the result line is already on stdout by then."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    killer = threading.Timer(20.0, lambda: os._exit(0))
    killer.daemon = True
    killer.start()
    tr._graph = None
    torch.cuda.synchronize()
    dist.destroy_process_group()
    killer.cancel()


if __name__ == "__main__":
    main()
