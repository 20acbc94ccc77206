#!/usr/bin/env python
"""bench.py -- clip-windows/sec of the JMT hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's own modules (oracle/_ref) on the host cores

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): one training step of
    TemporalConvNet(1024,[512]*4,k=5) on visual (B,1024,T)  +  FcLayer(768,512) on audio (B,T,768)
    -> Two_transformers(TRANSFORMER, FC, heads=1, layers=1) -> live CCC loss (V + A) -> backward -> SGD step
with B = 256 windows per GPU, T = 300, bf16 operands / fp32 accumulate, synthetic N(0,1) features,
random-init weights (seed 0).  A "step" is one pass over one batch; throughput = windows / second.
One process per GPU; the batch of windows is sharded across ranks (weak scaling: 256 per GPU), the flat
live-gradient bucket is all-reduced over NCCL at the end of backward, and the six CCC sums are all-reduced
inside the loss so every rank optimises the global-batch CCC.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GFLOP_PER_WINDOW = 0.236 + 10.15 + 7.23     # FcLayer + w_JR/FC + TCN useful taps, T=300
METRIC = "clip-windows/sec fwd+bwd"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="jmt", choices=["jmt", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="windows per GPU")
    ap.add_argument("--seq", type=int, default=300)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--heads", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=32, help="windows per step of the CPU baseline / reference arm")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-process parity record (pred_rel / ccc_delta)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record of a multi-GPU run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the captured step")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling (SURVEY C3): --batch is the GLOBAL number of windows, split evenly over the ranks")
    ap.add_argument("--gemm-table", default=None, help="write the per-shape GEMM timing table (roofline pass) to this file")
    return ap.parse_args()


# --------------------------------------------------------------------------------------- CPU arms
def cpu_port_step(B, T, heads, threads, seed=0):
    """One fwd+bwd of the same pipeline through the CPU oracle (the reference algorithm restated with plain
    torch CPU ops, fp32, all host threads).  Returns (seconds, windows).  Fallback when oracle/_ref is absent."""
    from oracle import jmt_oracle as O
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(seed)
    fus = {k: v.requires_grad_(True) for k, v in O.synth_params(O.two_transformers_shapes(1, "TRANSFORMER", "FC", 512, include_dead=False), 1).items()}
    fc = {k: v.requires_grad_(True) for k, v in O.synth_params([("fc_layer.weight", (512, 768)), ("fc_layer.bias", (512,))], 2).items()}
    tcn = {k: v.requires_grad_(True) for k, v in O.synth_params(O.tcn_shapes(1024, [512] * 4, 5), 3).items()}
    vis = torch.randn(B, 1024, T, generator=gen)
    aud = torch.randn(B, T, 768, generator=gen)
    lv, la = O.synth_labels(B, T, 4)
    t0 = time.perf_counter()
    vfeat = O.tcn_forward(vis, tcn, 4).transpose(1, 2)
    afeat = O.fc_layer_forward(aud, fc)
    v, a = O.two_transformers_forward(afeat, vfeat, fus, heads, 1, "TRANSFORMER", "FC")
    loss = O.ccc_loss_live(v, lv) + O.ccc_loss_live(a, la)
    loss.backward()
    return time.perf_counter() - t0, B


class ReferenceCpuStep:
    """The REFERENCE'S OWN modules (oracle/_ref, see oracle/make_ref.py) running the benchmarked training step on the host
    cores: TemporalConvNet(1024,[512]*4,k=5) -> transpose (I3DWSDDA.py:44) | FcLayer(768,512) -> Two_transformers(TRANSFORMER,
    FC) -> live CCC loss (train.py:283-311) -> backward -> SGD step, fp32, training mode, stock code path
    (nn.MultiheadAttention with need_weights=True and all)."""

    def __init__(self, heads, threads):
        from oracle import make_ref
        R = make_ref.import_reference()
        torch.set_num_threads(threads)
        torch.manual_seed(0)
        self.fusion = R["Two_transformers"](0.0, 0.0, heads, 1, "TRANSFORMER", "FC", 512).train()
        self.fc = R["FcLayer"](768, 512).train()
        self.tcn = R["TemporalConvNet"](1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1).train()
        self.crit = R["live_loss"]
        params = [p for m in (self.fusion, self.fc, self.tcn) for p in m.parameters()]
        self.opt = torch.optim.SGD(params, lr=1e-3)

    def __call__(self, B, T, seed=0):
        gen = torch.Generator().manual_seed(seed)
        vis = torch.randn(B, 1024, T, generator=gen)
        aud = torch.randn(B, T, 768, generator=gen)
        lv = torch.rand(B, T, generator=gen) * 2 - 1
        la = torch.rand(B, T, generator=gen) * 2 - 1
        n = B * T
        t0 = time.perf_counter()
        vfeat = self.tcn(vis).transpose(1, 2).contiguous()
        v, a = self.fusion(self.fc(aud), vfeat)
        loss = self.crit(v.reshape(-1, n), lv.reshape(-1, n)) + self.crit(a.reshape(-1, n), la.reshape(-1, n))
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return time.perf_counter() - t0, B


def make_cpu_arm(heads, threads):
    """(callable(B, T) -> (seconds, windows), kind, description)"""
    from oracle import make_ref
    if make_ref.available():
        return ReferenceCpuStep(heads, threads), "reference", "the reference's own modules from oracle/_ref (fp32, torch CPU)"
    return (lambda B, T: cpu_port_step(B, T, heads, threads)), "port", "oracle/jmt_oracle.py (restatement; oracle/_ref absent)"


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on the box's host cores, on my arm's config /
    metric / unit; each step is a bounded sample (--cpu-sample windows) of the 256-window workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = args.cpu_sample
    arm, kind, what = make_cpu_arm(args.heads, cores)
    for _ in range(args.warmup):
        arm(B, args.seq)
    times = []
    for _ in range(args.steps):
        dt, _ = arm(B, args.seq)
        times.append(dt)
        if sum(times) > 240:
            break
    med = float(np.median(times))
    val = B / med
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": kind,
                             "sample": f"{B} windows x T={args.seq} per step (bounded sample of the {args.batch}-window step), "
                                       f"fwd+bwd+SGD, fp32, {what}, median of {len(times)} steps"},
            "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "C2: TCN(1024,[512]*4,k=5)+FcLayer(768,512)+Two_transformers(TRANSFORMER,FC,h=%d,L=1)+CCC train step" % args.heads,
            "windows_per_gpu": args.batch, "seq_len": args.seq, "global_windows": args.batch * args.gpus,
            "parallelism": f"dp{args.gpus} (batch of windows sharded, NCCL grad all-reduce + CCC-sum all-reduce)",
            "optimizer": "SGD(lr=1e-3)", "l2": "inputs+activations per step (>1 GB) exceed the 126 MB L2; no explicit flush",
            "host_feature_dtype": "bf16"}


# --------------------------------------------------------------------------------------- parity record
def parity_record(precision, dev):
    """BASELINE.json metric, third part ("CCC delta vs ref"): the benchmarked pipeline at B = 4, T = 300 with default-init
    weights (seeded constructors reproduce the reference's initialisation bit for bit, tests/test_host_cpu.py) against the
    REFERENCE's outputs stored in tests/golden/piped_b4_t300.npz (tests/golden/make_golden.py).  No oracle involved."""
    import jmt_b200
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")))["piped_b4_t300"]
    g = np.load(os.path.join(ROOT, "tests", "golden", "piped_b4_t300.npz"))
    B, T = meta["B"], meta["T"]
    torch.manual_seed(meta["init_seed"])
    model = jmt_b200.JMTPipeline(jmt_b200.Two_transformers(0.0, 0.0, 1, 1, "TRANSFORMER", "FC", 512, precision=precision),
                                 jmt_b200.FcLayer(768, 512, precision=precision),
                                 jmt_b200.TemporalConvNet(1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1,
                                                          precision=precision)).to(dev).eval()
    gen = torch.Generator().manual_seed(meta["data_seed"])
    vis = torch.randn(B, 1024, T, generator=gen)
    aud = torch.randn(B, T, 768, generator=gen)
    with torch.no_grad():
        v, a = model(aud.to(dev), vis.to(dev))
    ref_v, ref_a = torch.from_numpy(g["vout"]).to(dev), torch.from_numpy(g["aout"]).to(dev)
    lv, la = torch.from_numpy(g["lv"]).to(dev), torch.from_numpy(g["la"]).to(dev)     # (B, T); predictions are (T, B): SURVEY Q1
    pred_rel = max(float((v - ref_v).abs().max() / ref_v.abs().max()), float((a - ref_a).abs().max() / ref_a.abs().max()))
    ccc = jmt_b200.cccmetric.ccc
    d_v = abs(ccc(v.reshape(-1), lv.reshape(-1)) - ccc(ref_v.reshape(-1), lv.reshape(-1)))     # flattened like train.py:303-307
    d_a = abs(ccc(a.reshape(-1), la.reshape(-1)) - ccc(ref_a.reshape(-1), la.reshape(-1)))
    crit = jmt_b200.CCCLoss(digitize_num=1)
    n = B * T
    loss = float((crit(v.view(-1, n), lv.view(-1, n)) + crit(a.view(-1, n), la.view(-1, n))).item())
    return {"case": "piped_b4_t300: the benchmarked pipeline at B=4, T=300, default init, vs the reference's outputs (tests/golden)",
            "precision": precision, "pred_rel": pred_rel, "ccc_delta": max(d_v, d_a), "loss_delta": abs(loss - float(g["loss"]))}


# --------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
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
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    import signal
    signal.alarm(1500)            # a wedged collective / teardown must end the process, not the GPU box's time limit
    if args.impl == "reference":
        run_reference(args)
        return
    import jmt_b200
    from jmt_b200 import engine as E
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = jmt_b200.dist.init_from_env("nccl") if world > 1 else 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.strong:
        if args.batch % world:
            raise SystemExit(f"--strong needs --batch ({args.batch}) divisible by the number of ranks ({world})")
        args.batch //= world
    B, T = args.batch, args.seq

    torch.manual_seed(0)                                   # identical weights on every rank
    fusion = jmt_b200.Two_transformers(0.0, 0.0, args.heads, 1, "TRANSFORMER", "FC", 512, precision=args.precision)
    fc = jmt_b200.FcLayer(768, 512, precision=args.precision)
    tcn = jmt_b200.TemporalConvNet(1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1, precision=args.precision)
    model = jmt_b200.JMTPipeline(fusion, fc, tcn).to(dev).train()
    crit = jmt_b200.CCCLoss(digitize_num=1, global_stats=world > 1)
    if world > 1:
        jmt_b200.dist.broadcast_parameters(model)
        # the loss is the GLOBAL-batch CCC (sums all-reduced in the forward): its parameter gradient is the SUM over ranks
        model.set_grad_sync(jmt_b200.dist.make_grad_sync(global_loss=True))
    opt = torch.optim.SGD(model.live_parameters(), lr=1e-3)

    # synthetic features, per-rank seed; host copies pinned (bf16 feature shards), device copies resident
    gen = torch.Generator().manual_seed(100 + rank)
    n_buf = 2
    host = []
    for i in range(n_buf):
        vis = torch.randn(B, 1024, T, generator=gen).to(torch.bfloat16).pin_memory()
        aud = torch.randn(B, T, 768, generator=gen).to(torch.bfloat16).pin_memory()
        lv = (torch.rand(B, T, generator=gen) * 2 - 1)
        la = (torch.rand(B, T, generator=gen) * 2 - 1)
        drop = torch.rand(B, T, generator=gen) < 0.05
        lv = torch.where(drop, torch.full_like(lv, -5.0), lv).pin_memory()
        la = la.pin_memory()
        host.append((aud, vis, lv, la))
    resident = [tuple(t.to(dev) for t in h) for h in host]
    n = B * T

    def step(aud, vis, lv, la):
        v, a = model(aud, vis)
        # train.py:303-311: v_loss + a_loss on the independently flattened predictions / labels; one fused call so that the
        # valence and arousal sums share ONE all-reduce across ranks
        loss = crit.forward_va(v.view(-1, n), lv.view(-1, n), a.view(-1, n), la.view(-1, n))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(*resident[i % n_buf])
    barrier()

    # host cost of one eagerly launched step (ctypes calls, tensor-map encodes, tape closures), without any sync
    t_h = time.perf_counter()
    for i in range(2):
        step(*resident[i % n_buf])
    host_enqueue_ms = (time.perf_counter() - t_h) / 2 * 1e3
    barrier()

    # the product path: the whole step (forward, CCC loss, backward, NCCL all-reduce, SGD) replayed as one CUDA graph
    # per input buffer set
    graphed = None
    if not args.no_graph:
        graphed = jmt_b200.GraphedStep(step, resident, warmup=3)
        run_step = lambda i: graphed.replay(i % n_buf)            # noqa: E731
        for i in range(3):
            run_step(i)
    else:
        run_step = lambda i: step(*resident[i % n_buf])           # noqa: E731
    barrier()

    # ---- timed region: K steps, inputs resident in HBM, CUDA events on the launching stream
    sampler = ClockSampler(local)
    sampler.start()
    l0 = jmt_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = run_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # kernels of ours executed in the timed region: counted at launch (eager) or at capture x replays (graph)
    launches = graphed.launches_per_step * args.steps if graphed is not None else jmt_b200.launch_count() - l0
    clocks = sampler.result()
    final_loss = float(loss.item())
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = B * world * args.steps / (ms / 1e3)

    # ---- e2e: same step through the public API with HOST (pinned) inputs; H2D of every step's inputs and
    #      a D2H read of the loss inside the timed region (copies double-buffered on a side stream)
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()
        # the captured graphs read the `resident` buffer sets: the H2D copies land there
        dbuf = resident if graphed is not None else [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i % 2])
                for d, h in zip(dbuf[i % 2], host[i % n_buf]):
                    d.copy_(h, non_blocking=True)
                ready[i % 2].record(copy_stream)

        h2d = sum(t.numel() * t.element_size() for t in host[0])
        for ev in consumed:
            ev.record()
        # every step's loss is copied to pinned host memory on a third stream as soon as that step is done and is read by
        # the host one step later, so the next step is already enqueued while the host waits (asynchronous logging; the
        # loss tensor of a buffer set is only overwritten two steps later)
        d2h_stream = torch.cuda.Stream()
        host_loss = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
        step_done = [torch.cuda.Event(), torch.cuda.Event()]
        d2h_done = [torch.cuda.Event(), torch.cuda.Event()]
        barrier()
        t0 = time.perf_counter()
        prefetch(0)
        lsum = 0.0
        for i in range(args.steps):
            if i + 1 < args.steps:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            l = graphed.replay(i % 2) if graphed is not None else step(*dbuf[i % 2])
            consumed[i % 2].record()
            step_done[i % 2].record()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(step_done[i % 2])
                host_loss[i:i + 1].copy_(l.detach().reshape(1).float(), non_blocking=True)   # D2H of the step's result
                d2h_done[i % 2].record(d2h_stream)
            if i >= 1:
                d2h_done[(i - 1) % 2].synchronize()
                lsum += float(host_loss[i - 1])
        d2h_done[(args.steps - 1) % 2].synchronize()
        lsum += float(host_loss[args.steps - 1])
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * args.steps / float(tt.item()), "unit": "windows/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4}

    # ---- roofline of the dominant kernel (gemm_tc_kernel): per-launch CUDA events over timed steps
    roof = None
    if not args.no_roofline:              # every rank runs the profiled steps (they contain collectives)
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        E.PROFILE = []
        nprof = min(args.steps, 3)
        for i in range(nprof):
            step(*resident[i % n_buf])
        torch.cuda.synchronize()
        rec, E.PROFILE = E.PROFILE, None
        kname = "gemm_simt_kernel" if args.precision == "fp32" else "gemm_tc_kernel"
        sel = [r for r in rec if r[0] == kname]
        tot_ms = sum(r[2].elapsed_time(r[3]) for r in sel)
        tot_fl = sum(r[1] for r in sel)
        ach = tot_fl / (tot_ms / 1e3) / 1e12 if tot_ms > 0 else 0.0
        att = [r for r in rec if r[0] == "attn_chain_kernel"]
        att_ms = sum(r[2].elapsed_time(r[3]) for r in att)
        att_fl = sum(r[1] for r in att)
        dq = [r for r in rec if r[0] == "attn_bwd_dqkv_kernel"]        # dQ / dK / dV of the attention backward (DRAM-bound)
        dq_ms = sum(r[2].elapsed_time(r[3]) for r in dq)
        dq_fl = sum(r[1] for r in dq)
        # DRAM traffic of the dominant kernel's most frequent launch (76800x512x512 linear, 40 of 147 GEMM launches per step)
        # from the committed `ncu --set full` capture; algorithmic bytes of that launch = A + W + D = 157.8 MB
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_gemm_tc_r3p_linear_76800x512x512.csv")
        if kname == "gemm_tc_kernel" and os.path.exists(tpath):
            import csv
            rows = {r[0]: r for r in csv.reader(open(tpath)) if len(r) == 3}
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            try:
                traffic = sum(float(rows[k][2]) * mult[rows[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                traffic_src = "profiles/ncu_gemm_tc_r3p_linear_76800x512x512.csv (dram read+write of one 76800x512x512 launch; algorithmic 157.8e6 B)"
            except (KeyError, ValueError):
                traffic = None
        roof = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)",
                "launches_per_step": len(sel) // nprof, "gemm_ms_per_step": tot_ms / nprof,
                "gemm_gflop_per_step": tot_fl / nprof / 1e9,
                "avg_launch_us": 1e3 * tot_ms / max(1, len(sel)),
                "step_model_gflop": 3 * FWD_GFLOP_PER_WINDOW * B,
                "attn_chain_kernel": {"launches_per_step": len(att) // nprof, "ms_per_step": att_ms / nprof,
                                      "gflop_per_step": att_fl / nprof / 1e9,
                                      "achieved_tflops": att_fl / (att_ms / 1e3) / 1e12 if att_ms > 0 else 0.0},
                "attn_bwd_dqkv_kernel": {"launches_per_step": len(dq) // nprof, "ms_per_step": dq_ms / nprof,
                                         "gflop_per_step": dq_fl / nprof / 1e9,
                                         "achieved_tflops": dq_fl / (dq_ms / 1e3) / 1e12 if dq_ms > 0 else 0.0,
                                         "bound": "hbm", "dram_mb_per_launch_algorithmic": 564.0},
                "tensor_kernels_tflops": ((tot_fl + att_fl + dq_fl) / ((tot_ms + att_ms + dq_ms) / 1e3) / 1e12
                                          if tot_ms + att_ms + dq_ms > 0 else 0.0)}
        if args.gemm_table and rank == 0:
            agg = {}
            for r in rec:
                key = (r[0],) + tuple(r[4])
                a = agg.setdefault(key, [0, 0.0, 0.0])
                a[0] += 1
                a[1] += r[2].elapsed_time(r[3])
                a[2] += r[1]
            with open(args.gemm_table, "w") as f:
                f.write("kernel M N K nb taps | launches/step ms/step GFLOP/step TFLOP/s\n")
                for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    f.write(f"{key[0]} {key[1]} {key[2]} {key[3]} {key[4]} {key[5]} | {a[0] / nprof:.1f} {a[1] / nprof:.3f} "
                            f"{a[2] / nprof / 1e9:.1f} {a[2] / (a[1] / 1e3) / 1e12 if a[1] > 0 else 0:.1f}\n")
        if world > 1:
            dist.barrier()

    # ---- strong scaling: the SAME global batch of args.batch windows split over the ranks, K replayed steps
    strong = None
    if world > 1 and not args.strong and not args.no_strong and args.batch % world == 0 and not args.no_graph:
        Bs = args.batch // world
        ns = Bs * T
        sres = [tuple(t[:Bs].contiguous() for t in r) for r in resident]

        def sstep(aud, vis, lv, la):
            v, a = model(aud, vis)
            loss = crit.forward_va(v.view(-1, ns), lv.view(-1, ns), a.view(-1, ns), la.view(-1, ns))
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss
        sg = jmt_b200.GraphedStep(sstep, sres, warmup=3)
        for i in range(3):
            sg.replay(i % n_buf)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(args.steps):
            sg.replay(i % n_buf)
        s1.record()
        barrier()
        sms = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        strong = {"global_windows": args.batch, "windows_per_gpu": Bs, "ms_per_step": float(sms.item()) / args.steps,
                  "value": args.batch * args.steps / (float(sms.item()) / 1e3), "unit": "windows/s", "scaling": "strong",
                  "note": "same global batch as the 1-GPU run split over the ranks; CUDA-graph replay, device-timed, max over ranks"}
        sg.graphs.clear()
        sg.losses.clear()

    # ---- parity of the benchmarked precision (and of the bf16x3 gate mode) against the reference's outputs
    parity = None
    if rank == 0 and not args.no_parity:
        parity = {args.precision: parity_record(args.precision, dev)}
        if args.precision != "bf16x3":
            parity["bf16x3"] = parity_record("bf16x3", dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        arm, kind, what = make_cpu_arm(args.heads, cores)
        arm(2, T)                                              # warm-up (thread pools, allocator)
        times = []
        t_start = time.perf_counter()
        while len(times) < 5 and time.perf_counter() - t_start < 30:
            dt, _ = arm(args.cpu_sample, T)
            times.append(dt)
        med = float(np.median(times))
        cpu = {"value": args.cpu_sample / med, "unit": "windows/s", "cores": cores, "kind": kind,
               "sample": f"{args.cpu_sample} windows x T={T} per step, fwd+bwd+SGD, fp32, median of {len(times)} steps of {what}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (fp32 values as two bf16 parts, three tcgen05.mma per k-step)", "fp32": "f32"}[args.precision], "data": "synthetic",
                "config": workload_config(args), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu, "final_loss": final_loss,
                "ccc_delta": parity[args.precision]["ccc_delta"] if parity else None,
                "pred_rel": parity[args.precision]["pred_rel"] if parity else None, "parity": parity, "strong": strong,
                "launch_mode": "cuda_graph" if graphed is not None else "eager", "host_enqueue_ms_per_eager_step": host_enqueue_ms,
                "peak_hbm_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
                "model_tflops": 3 * FWD_GFLOP_PER_WINDOW * value / 1e3}
        print(json.dumps(line), flush=True)
    # Teardown: captured graphs hold NCCL kernels; destroying the process group under them hung a 2-GPU run once
    # (after the JSON line was out).  Drop the graphs, drain the device, meet at a barrier and leave without the
    # NCCL destructor -- the OS reclaims the communicators.
    if graphed is not None:
        graphed.graphs.clear()
        graphed.losses.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
