#!/usr/bin/env python
"""bench.py -- fusion-head forward+backward throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One step = one forward + backward of the fusion head over one batch of synthetic encoder features
(random-init weights), per-GPU batch fixed (weak scaling); N>1 adds the all-gather inside InfoNCE
and one gradient all-reduce.  Prints ONE JSON line (rank 0) with the device-timed `value`, the end-to-end
`e2e` (pinned host inputs -> H2D -> step -> D2H of the loss, through the public nn.Module API), the `roofline`
of the dominant kernel (tcgen05 GEMM, timed live with CUDA events) and the `cpu_baseline` (oracle port on the
host cores, bounded sample).  `--impl reference` times the reference's CPU path (oracle port -- the reference
is pure Python/torch and cannot travel to the GPU box) on the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch                                                             # noqa: E402
import torch.distributed as dist                                         # noqa: E402

# name -> (head kind, per-GPU batch, (Lt, La, Lv) or None for 2-D, contrastive flag, CPU sample batch)
WORKLOADS = {
    "mult_b256": ("mult", 256, (512, 512, 30), False, 4),                      # BASELINE configs[1]
    "hierarchical_b4096": ("hierarchical", 4096, (512, 512, 30), True, 4),     # BASELINE configs[3] / north_star target
    "contrastive_b4096": ("contrastive", 4096, None, True, 4096),              # BASELINE configs[2]
    "early_b16": ("early", 16, None, False, 16),                               # BASELINE configs[0]
}
DEFAULT_WORKLOAD = "hierarchical_b4096"          # the north_star target configuration (BASELINE.json configs[3])
SECONDARY = ("mult_b256", "contrastive_b4096")     # also measured (value + e2e only) and reported under `secondary`
H, HEADS = 512, 8
# algorithmic forward GFLOP per sample (2*MACs of the reference's contractions only); fwd+bwd = 3x
MULT_GFLOP = 17.7495


def algorithmic_gflop_per_sample(kind, lens, b_global):
    if kind == "mult":
        return 3 * MULT_GFLOP
    if kind == "hierarchical":
        small = (4.19 + 19.40 + 3.93 + 9.98 + 6.29) * 1e-3 + 3 * 2 * b_global * 256 * 1e-9
        return 3 * (MULT_GFLOP + small)
    if kind == "contrastive":
        return 3 * (3.93e-3 + 3 * 2 * b_global * 256 * 1e-9)
    if kind == "early":
        return 3 * 4.19e-3
    raise ValueError(kind)


class Cfg:
    fusion_hidden_size, fusion_num_heads, fusion_dropout = H, HEADS, 0.0
    num_emotions, graph_hidden_size, graph_num_layers, graph_dropout = 7, 512, 3, 0.0
    contrastive_temperature = 0.07


def gemm_traffic():
    """DRAM bytes of one launch of the dominant kernel, from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm": d.get("hbm_gbs"), "src": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """start of the window the summary is taken over (the sampler itself is started earlier so that it is warm)"""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time()
        time.sleep(0.15)
        self.proc.terminate()
        t0 = getattr(self, "t_mark", 0.0)
        rows = [r for t, r in self.rows if t0 <= t <= t_end + 0.12] or [r for _, r in self.rows[-3:]]
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def objective(out, b_global):
    """mean(fused^2) over the GLOBAL batch + 0.1 * sum(InfoNCE) -- a [B,512] reduction, negligible work."""
    t = out if isinstance(out, torch.Tensor) else out["fused_features"]
    loss = (t.float() ** 2).sum() / (b_global * t.size(-1))
    if isinstance(out, dict):
        for v in out.get("contrastive_losses", {}).values():
            loss = loss + 0.1 * v
    return loss


# ------------------------------------------------------------------------------------------------------------
def run_reference(args, kind, lens, flag, cpu_batch, world, rank):
    """The reference's CPU path (oracle port: same torch CPU operators the reference's modules dispatch to)."""
    if rank != 0:
        return
    from oracle import fusion_oracle as fo
    fo.TRAIN_DROPOUT = args.dropout
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    P = {k: v.requires_grad_(True) for k, v in fo.init_params(kind, H=H, heads=HEADS, seed=0).items()}
    feats = [f.requires_grad_(True) for f in fo.synthetic_features(cpu_batch, lens or (None, None, None), H=H, seed=1234)]
    kw = {}
    if kind in ("contrastive", "hierarchical"):
        kw["compute_contrastive_loss"] = flag

    def step():
        for t in list(P.values()) + feats:
            t.grad = None
        out = fo.HEADS[kind](*feats, P, **kw)
        loss = objective(out, cpu_batch)
        loss.backward()
        return float(loss.detach())

    for _ in range(max(args.warmup, 1)):             # at least one untimed step: MKL / allocator first-use costs
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = cpu_batch / dt
    line = {"impl": "reference", "metric": "fusion_head_fwd_bwd_samples_per_sec", "value": val, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "head": kind, "seq_lens": lens, "hidden": H, "dropout": args.dropout, "note": "CPU, bounded sample"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": f"{cpu_batch} samples/step x {args.steps} steps of the same workload (fp32, torch CPU ops, all host threads)"},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def cpu_baseline(kind, lens, flag, cpu_batch, budget_s=20.0, dropout=0.0):
    from oracle import fusion_oracle as fo
    fo.TRAIN_DROPOUT = dropout
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    P = {k: v.requires_grad_(True) for k, v in fo.init_params(kind, H=H, heads=HEADS, seed=0).items()}
    feats = [f.requires_grad_(True) for f in fo.synthetic_features(cpu_batch, lens or (None, None, None), H=H, seed=1234)]
    kw = {"compute_contrastive_loss": flag} if kind in ("contrastive", "hierarchical") else {}
    times = []
    t_start = time.perf_counter()
    while len(times) < 4 and (time.perf_counter() - t_start < budget_s or len(times) < 2):
        for t in list(P.values()) + feats:
            t.grad = None
        t0 = time.perf_counter()
        objective(fo.HEADS[kind](*feats, P, **kw), cpu_batch).backward()
        times.append(time.perf_counter() - t0)
    dt = statistics.median(times[1:]) if len(times) > 1 else times[0]
    return {"value": cpu_batch / dt, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{cpu_batch} samples/step, median of {max(1, len(times) - 1)} steps after 1 warm-up (oracle port, fp32, torch CPU ops)"}


# ------------------------------------------------------------------------------------------------------------
_RESULT_FD = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: native libraries write there too (NCCL prints its version banner to stdout when the box
    sets NCCL_DEBUG=VERSION), so fd 1 is pointed at stderr for the run and the result goes to a duplicate of the original fd 1."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


class Runner:
    """One workload on this rank: builds the head and the synthetic features, picks how the step is issued and offers the
    timed legs (`device_leg`, `e2e_leg`, `gemm_profile_leg`)."""

    def __init__(self, args, name, world, rank, dev):
        import simple_multimodal_b200 as pkg
        self.pkg, self.args, self.name, self.world, self.rank, self.dev = pkg, args, name, world, rank, dev
        kind, batch, lens, flag, cpu_batch = WORKLOADS[name]
        if args.batch and name == args.workload:
            batch = args.batch
        self.kind, self.batch, self.lens, self.flag, self.cpu_batch = kind, batch, lens, flag, cpu_batch
        self.b_global = batch * world
        torch.manual_seed(0)                              # random-init weights, the reference's default initialisers
        Cfg.fusion_dropout = Cfg.graph_dropout = args.dropout
        head = getattr(pkg.fusion_layers, {"mult": "MultimodalTransformer", "hierarchical": "HierarchicalFusion",
                                           "contrastive": "ContrastiveFusion", "early": "EarlyFusion"}[kind])(Cfg).to(dev)
        head.train()
        self.head = head
        self.mult = head.mult_fusion if kind == "hierarchical" else (head if kind == "mult" else None)
        if args.chunk and self.mult is not None:
            self.mult.chunk_size = args.chunk
        if args.no_graph and self.mult is not None:
            self.mult.graph_chunks = False
        self.params = [p for p in head.parameters()]
        # every parameter gradient in one flat fp32 buffer (p.grad are views): zeroed by one fill, all-reduced in place
        self.bucket = pkg.GradBucket(self.params)
        # synthetic encoder features: pinned host buffers (e2e source) + resident device copies (device-timed `value`)
        g = torch.Generator().manual_seed(1234 + rank)
        shapes = [(batch, H)] * 3 if lens is None else [(batch, L, H) for L in lens]
        self.host = [torch.randn(s, generator=g).to(torch.bfloat16).pin_memory() for s in shapes]
        self.resident = [h.to(dev, non_blocking=True).requires_grad_(True) for h in self.host]
        # modality dropout (reference models/encoders.py:280-321): a fresh keep-mask every step, generated on the device
        self.md = pkg.ModalityDropout(0.1, seed=4321 + rank) if kind == "hierarchical" else None
        self.h2d_bytes = sum(h.numel() * h.element_size() for h in self.host)
        self.kw = {"compute_contrastive_loss": flag} if kind in ("contrastive", "hierarchical") else {}
        self.graphed, self.graph_inputs = None, None
        self.issue = "eager (--no-graph)"
        self._pick_issue_mode()

    def loss_of(self, out):
        return objective(out, self.b_global)

    def eager_step(self, inputs):
        self.bucket.zero()
        for x in inputs:
            x.grad = None
        kw = self.kw if self.md is None else dict(self.kw, mask=self.md.sample_mask(self.batch, self.dev))
        loss = self.loss_of(self.head(*inputs, **kw))
        loss.backward()
        self.bucket.all_reduce()
        return loss

    def _pick_issue_mode(self):
        """How a step reaches the GPU.  (1) Big MulT batches (more than one chunk): `ChunkGraphEngine` inside the head replays
        one captured graph per MulT chunk and direction; the small heads, the InfoNCE collectives and the gradient all-reduce
        are issued eagerly around them.  (2) Steps without a collective inside forward/backward: the whole step is ONE captured
        graph (`GraphedTrainStep`).  (3) Otherwise eager."""
        args, kind, world = self.args, self.kind, self.world
        if args.no_graph:
            return
        multi_chunk = self.mult is not None and self.lens is not None and self.batch > self.mult.chunk_size
        if multi_chunk:
            self.issue = "per-chunk CUDA graphs inside the head (ChunkGraphEngine); small heads / collectives issued eagerly"
            return
        if world > 1 and kind in ("contrastive", "hierarchical") and os.environ.get("B200F_GRAPH_COLLECTIVES", "0") != "1":
            # (capturing the step's NCCL collectives into the graph is opt-in: B200F_GRAPH_COLLECTIVES=1)
            self.issue = "eager (NCCL all-gather inside the step)"
            return
        try:
            if self.mult is not None:
                self.mult.graph_chunks = False            # the whole step is captured instead
            for _ in range(2):
                self.eager_step(self.resident)
            torch.cuda.synchronize()
            self.graphed = self.pkg.GraphedTrainStep(self.head, self.resident, self.loss_of, self.kw, grad_bucket=self.bucket,
                                                     modality_dropout=self.md)
            self.issue = "cuda-graph replay of forward+loss+backward (GraphedTrainStep)"
            self.graph_inputs = self.graphed.static_inputs    # the device-timed leg feeds the graph's own input buffers
        except Exception as exc:                              # noqa: BLE001 -- never let the capture take the bench down
            self.graphed, self.issue = None, f"eager (capture failed: {type(exc).__name__}: {str(exc)[:120]})"
            torch.cuda.synchronize()

    def step(self, inputs):
        if self.graphed is None:
            return self.eager_step(inputs)
        loss = self.graphed(*(self.graph_inputs if inputs is self.resident else inputs))
        self.bucket.all_reduce()
        return loss

    def sync_all(self):
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches_per_step(self, counted):
        """kernels of this library per step: those issued through the C ABI during the step + those inside replayed graphs"""
        if self.graphed is not None:
            return self.graphed.kernel_launches
        eng = getattr(self.mult, "_engine", None) if self.mult is not None else None
        return counted + (eng.kernel_launches if eng is not None else 0)

    def warmup(self, sampler=None):
        """the W requested steps, then keep stepping (bounded) until the clocks have ramped and the step time is stable -- the
        first process on a fresh box measured 36 ms/step for its first ~0.5 s and 22 ms/step afterwards"""
        args, warm_done, last, t_warm = self.args, 0, None, time.perf_counter()
        while True:
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            self.step(self.resident)
            w1.record()
            torch.cuda.synchronize()
            warm_done += 1
            cur = w0.elapsed_time(w1)
            stable = last is not None and abs(cur - last) <= 0.03 * last
            last = cur
            if warm_done == args.warmup:                    # the clock for the extra steps starts after the requested ones (lazy
                t_warm = time.perf_counter()                # initialisation, graph capture, NCCL set-up sit in the first steps)
            elapsed = time.perf_counter() - t_warm
            done = warm_done >= args.warmup and (args.fixed_warmup or (stable and elapsed >= 2.0) or elapsed >= 6.0 or warm_done >= args.warmup + 200)
            if self.world > 1:                              # steps contain collectives: every rank must run the same number of them
                flag = torch.tensor([0 if done else 1], device=self.dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                done = int(flag.item()) == 0
            if done:
                break
        self.sync_all()
        return warm_done

    def device_leg(self, steps, sampler=None):
        """K clean steps with the inputs resident in HBM -> (ms per step, host ms to issue a step, library launches per step)"""
        l0 = self.pkg._lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.sync_all()
        if sampler is not None:
            sampler.mark()
        e0.record()
        t_host, c_host = time.perf_counter(), time.process_time()
        for _ in range(steps):
            self.step(self.resident)
        host_issue_ms = (time.perf_counter() - t_host) * 1e3 / steps     # wall time the host spent inside the issue loop (no sync inside):
        self.host_cpu_ms = (time.process_time() - c_host) * 1e3 / steps  # it includes waiting on a full launch queue; the CPU time next to
        e1.record()                                                      # it is what the host really works per step -- if THAT approaches
        self.sync_all()                                                  # ms_per_step the run is launch-bound on the host, not GPU-bound
        counted = (self.pkg._lib.launch_count() - l0) // steps
        return e0.elapsed_time(e1) / steps, host_issue_ms, self.launches_per_step(counted)

    def e2e_leg(self, steps):
        """pinned host -> H2D -> step -> D2H of the loss, through the nn.Module API.  Every step's inputs are copied from pinned
        host memory inside the timed region; the copy of step k+1 is staged on a side stream (pkg.FeaturePrefetcher) while step k
        computes.  The loss of every step is read back inside the timed region one step late: step k's loss goes to a pinned word
        with an asynchronous D2H copy and is consumed after step k+1 has been issued (the way a training loop logs it)."""
        pf = self.pkg.FeaturePrefetcher(self.dev)
        loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ready = [torch.cuda.Event() for _ in range(2)]
        losses = []

        def run(n):
            pf.submit(self.host)
            for k in range(n):
                xs = pf.get() if self.graphed is not None else [x.requires_grad_(True) for x in pf.get()]
                if k + 1 < n:
                    pf.submit(self.host)
                loss_host[k % 2].copy_(self.step(xs).detach().float(), non_blocking=True)
                loss_ready[k % 2].record()
                if k > 0:
                    loss_ready[(k - 1) % 2].synchronize()
                    losses.append(float(loss_host[(k - 1) % 2]))
            loss_ready[(n - 1) % 2].synchronize()
            losses.append(float(loss_host[(n - 1) % 2]))

        run(2)
        self.sync_all()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        run(steps)
        t1.record()
        self.sync_all()
        if not all(l == l and abs(l) < 1e30 for l in losses):
            raise RuntimeError(f"non-finite loss in the e2e leg: {losses[-3:]}")
        return t0.elapsed_time(t1) / steps, losses[-1]

    def gemm_profile_leg(self, steps):
        """the same K steps again, issued eagerly with a CUDA-event pair around every GEMM launch (the dominant kernel family)
        -> `roofline`.  A separate pass: the event records cost launch gaps which must not sit in `value`; events cannot be
        recorded inside a graph replay, so the chunk graphs are released first (their pool holds the activation stash)."""
        from simple_multimodal_b200 import kernels as K
        if self.mult is not None:
            self.mult.graph_chunks = False
            self.mult.release_graphs()
            torch.cuda.empty_cache()
        K.prealloc_profile_events(2 * 100 * max(1, self.batch // 256 + 1) * steps + 2048)
        self.profile_chunk = None if self.mult is None else int(self.mult.chunk_size)
        try:
            self.eager_step(self.resident)                # untimed: the eager path's allocations happen here
        except torch.OutOfMemoryError:                    # the eager path keeps its temporaries outside any graph pool: retry with smaller chunks
            if self.mult is None:
                raise
            for p in self.params:
                p.grad = None
            torch.cuda.empty_cache()
            self.mult.chunk_size = self.profile_chunk = 256
            self.bucket.attach()
            self.eager_step(self.resident)
        K.GEMM_PROFILE = []
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.sync_all()
        p0.record()
        for _ in range(steps):
            self.eager_step(self.resident)
        p1.record()
        self.sync_all()
        prof, K.GEMM_PROFILE = K.GEMM_PROFILE, None
        gemm_ms = sum(a.elapsed_time(b) for a, b, _, tc in prof if tc) / steps
        gemm_flops = sum(f for _, _, f, tc in prof if tc) / steps
        n_gemm = sum(1 for *_, tc in prof if tc) // steps
        return p0.elapsed_time(p1) / steps, gemm_ms, gemm_flops, n_gemm

    def close(self):
        if self.mult is not None:
            self.mult.release_graphs()
        self.graphed = None
        self.head = self.bucket = self.params = self.resident = self.host = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()


def reduce_max(vals, world, dev):
    tt = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return [float(v) for v in tt]


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("B200F_WORKLOAD", DEFAULT_WORKLOAD), choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--chunk", type=int, default=0, help="MulT chunk size (samples)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads (mult_b256, contrastive_b4096)")
    ap.add_argument("--no-roofline-pass", action="store_true", help="skip the instrumented GEMM pass (profiler runs)")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from Python instead of replaying captured graphs")
    ap.add_argument("--fixed-warmup", action="store_true", help="run exactly --warmup warm-up steps (profiler runs: ncu serialises every launch)")
    ap.add_argument("--dropout", type=float, default=0.1,
                    help="fusion_dropout / graph_dropout of the head in training mode (reference default, config.py:30,41: 0.1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    kind, batch, lens, flag, cpu_batch = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, kind, lens, flag, cpu_batch, world, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the b200 arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_cores = []
    if os.environ.get("B200F_NO_NUMA_BIND", "") != "1":
        # before any pinned allocation: the process (and its staging buffers) onto the GPU's own NUMA node
        host_cores = importlib.import_module("simple-multimodal_b200").bind_host_to_gpu(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    t_start = time.perf_counter()
    run = Runner(args, args.workload, world, rank, dev)
    batch, b_global = run.batch, run.b_global
    sampler = ClockSampler(local_rank)
    sampler.start()
    warm_done = run.warmup()
    ms, host_issue_ms, launches = run.device_leg(args.steps, sampler)
    clocks = sampler.stop()
    ms_e2e, last_loss = run.e2e_leg(args.steps)
    if args.no_roofline_pass:
        ms_prof, gemm_ms, gemm_flops, n_gemm = 0.0, 0.0, 0.0, 0
    else:
        ms_prof, gemm_ms, gemm_flops, n_gemm = run.gemm_profile_leg(max(1, min(args.steps, 5)))
    ms, ms_e2e = reduce_max([ms, ms_e2e], world, dev)
    issue, h2d_bytes, host_cpu_ms = run.issue, run.h2d_bytes, getattr(run, "host_cpu_ms", None)
    profile_chunk = getattr(run, "profile_chunk", None)
    run.close()

    # ---- the other GPU-sized BASELINE configs, value + e2e only (never allowed to take the primary line down)
    secondary = {}
    if args.workload == DEFAULT_WORKLOAD and not args.no_secondary and not args.fixed_warmup:
        for name in SECONDARY:
            try:
                r2 = Runner(args, name, world, rank, dev)
                r2.warmup()
                ms2, host2, l2 = r2.device_leg(args.steps)
                e2e2, _ = r2.e2e_leg(args.steps)
                ms2, e2e2 = reduce_max([ms2, e2e2], world, dev)
                k2, _, lens2, _, _ = WORKLOADS[name]
                gf2 = algorithmic_gflop_per_sample(k2, lens2, r2.b_global)
                secondary[name] = {"value": r2.b_global / (ms2 * 1e-3), "unit": "samples/s", "ms_per_step": ms2, "batch_per_gpu": r2.batch,
                                   "e2e": r2.b_global / (e2e2 * 1e-3), "e2e_ms_per_step": e2e2, "host_issue_ms_per_step": host2,
                                   "gpu_launches": int(l2), "issue": r2.issue, "model_tflops_per_gpu": gf2 * r2.batch / 1e3 / (ms2 * 1e-3)}
                r2.close()
            except Exception as exc:                          # noqa: BLE001
                secondary[name] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
                if world > 1:                                 # a rank that failed alone would leave the others in a collective
                    raise

    if rank == 0:
        pk = peaks()
        traffic = gemm_traffic()
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        gflop = algorithmic_gflop_per_sample(kind, lens, b_global)
        model_tflops = gflop * batch / 1e3 / (ms * 1e-3)
        line = {
            "metric": "fusion_head_fwd_bwd_samples_per_sec", "value": b_global / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "head": kind, "batch_per_gpu": batch, "global_batch": b_global, "seq_lens": lens,
                       "hidden": H, "heads": HEADS, "dropout": args.dropout, "modality_dropout": 0.1 if kind == "hierarchical" else None,
                       "parallelism": f"dp{world}", "warmup_steps_run": warm_done, "host_issue_ms_per_step": host_issue_ms, "host_cpu_ms_per_step": host_cpu_ms, "host_cores_bound": len(host_cores), "issue": issue,
                       "collectives_per_step": (["all_gather(z)", "all_reduce(loss)", "all_gather(lse)", "all_reduce(grads, flat fp32 bucket)"]
                                                if world > 1 and kind in ("contrastive", "hierarchical") else
                                                (["all_reduce(grads, flat fp32 bucket)"] if world > 1 else [])),
                       "l2": "inputs+activations per step exceed the 126 MB L2" if h2d_bytes > 126e6 else "small working set (latency-bound config)",
                       "algorithmic_tflop_per_step": gflop * batch / 1e3,
                       "model_tflops_per_gpu": model_tflops,
                       "frac_of_sustained_peak": model_tflops / pk["tflops"] if pk["tflops"] else None,
                       "frac_of_nominal_2250": model_tflops / 2250.0,
                       "last_loss": last_loss, "wall_s": time.perf_counter() - t_start},
            "e2e": {"value": b_global / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_pair_kernel / gemm_tc_kernel (tcgen05 bf16 GEMM, all layouts)", "achieved": achieved,
                         "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"] if pk["tflops"] else None,
                         "traffic": (traffic or {}).get("traffic_bytes_per_launch"),
                         "traffic_launch": None if traffic is None else f'{traffic["kernel"]}: {traffic["launch"]}; algorithmic bytes '
                                                                          f'{traffic["algorithmic_bytes_per_launch"]} ({traffic["source"]})',
                         "peak_source": pk["src"], "launches_per_step": n_gemm, "kernel_ms_per_step": gemm_ms,
                         "share_of_step": gemm_ms / ms_prof if ms_prof else None, "instrumented_ms_per_step": ms_prof,
                         "note": "events around every tcgen05 GEMM launch in a separate eagerly issued pass of the same step"
                                 + (f" (MulT chunk {profile_chunk})" if profile_chunk else "")},
        }
        if secondary:
            line["secondary"] = secondary
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(kind, lens, flag, cpu_batch, dropout=args.dropout)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
