#!/usr/bin/env python
"""Benchmark of the MinGraph-UNet graph block on B200 (driver contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg1|cfg5]

One "step" = one pass of the hot path over one batch of synthetic images:
  per-pixel feature map (B,20,H,W) -> patch-mean pool -> 4-connected patch grid -> patch GAT ->
  predictor GAT -> softmax/argmax -> N-cut loss -> region mean-pool -> region GAT -> nearest
  un-pool written straight into the channel slice [32:96] of a (B,96,H,W) fusion buffer.

Headline line (every N): BASELINE.json configs[1] per GPU (512x512, batch 16 per GPU, bf16 storage, fp32 math), i.e.
weak scaling by image.  Ranks exchange only the small per-image outputs (loss, region features, labels): the block
kernel itself stores them into every rank's gathered buffer over NVLink peer memory (distributed.PeerExchange; no
collective call in the step; --exchange inline = NCCL all-gather fallback).
The same invocation also measures, as `extra` keys of the same JSON line, the two multi-GPU configs BASELINE.json names:
  extra.cfg3  configs[2]: 1024x1024, GLOBAL batch 64 sharded by image over the N GPUs (strong scaling; efficiency against
              the whole batch on rank 0 alone, measured in the same run);
  extra.cfg5  configs[4]: training step (forward + backward scatter + NCCL gradient all-reduce + Adam) at 512x512,
              4 images per GPU (= batch 32 on 8 GPUs; weak scaling; efficiency against rank 0 alone without all-reduce).
  extra.cfg4  configs[3]: a few points of the graph-layer micro-benchmark sweep (N = 262 144 nodes, k = 8 / 32, F = 64 / 128,
              4 heads, bf16; one GAT layer through the C ABI, L2 flushed before every timed launch) on rank 0: ms, edges/s,
              SpMM gather-model GB/s and its fraction of the measured HBM peak (the full sweep: tools/sweep.py).

Prints ONE JSON line (rank 0).  ``--impl reference`` times the UNTOUCHED reference classes on the host cores
(oracle/_ref, copied there by __graft_entry__.build(); falls back to the oracle port oracle/restate.py, labelled, when
the copy is absent).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (H, W, images per GPU, storage dtype)
    "cfg1": (256, 256, 1, "float32"),      # BASELINE configs[0]: the reference's CPU-runnable case
    "cfg2": (512, 512, 16, "bfloat16"),    # BASELINE configs[1]: headline single-GPU config
    "cfg3": (1024, 1024, 8, "bfloat16"),   # BASELINE configs[2] shard: 64 images over 8 GPUs
    "cfg5": (512, 512, 4, "bfloat16"),     # BASELINE configs[4] shard: TRAINING step, 32 images over 8 GPUs
}
CFG3_GLOBAL_BATCH = 64
IN_DIM, D_OUT, HEADS, K_SEG, PATCH, C_UNET = 20, 64, 4, 2, 16, 32
FALLBACK_HBM_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--shards", type=int, default=1, help="parallel sub-batches inside the captured graph (ours arm)")
    ap.add_argument("--pipeline-depth", type=int, default=3,
                    help="independent steps in flight (PipelinedGraphBlock slots; 1 = one replay at a time)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "inline"],
                    help="N>1: p2p = payload pushed by the block kernel over NVLink peer memory (default); "
                         "inline = one NCCL all-gather per step (fallback)")
    ap.add_argument("--no-numa-bind", dest="numa_bind", action="store_false",
                    help="N>1: do not bind each rank to its GPU's NUMA node (A/B of the e2e leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra.cfg3 / extra.cfg5 legs (profiling runs)")
    ap.add_argument("--cpu-sample-images", type=int, default=0, help="images in the CPU baseline sample (0 = auto)")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_config(name, n_gpus, B=None):
    H, W, B0, dt = WORKLOADS[name]
    B = B0 if B is None else B
    nph, npw = -(-H // PATCH), -(-W // PATCH)
    N = nph * npw
    E = 2 * (nph * (npw - 1) + npw * (nph - 1))
    return dict(workload=f"{name}: {H}x{W} images, batch {B}/GPU, {dt} storage + fp32 math, graph block "
                         f"(pool->grid graph->patch GAT->N-cut->region GAT->un-pool into fusion buffer)",
                H=H, W=W, images_per_gpu=B, global_batch=B * n_gpus, nodes_per_image=N, edges_per_image=E,
                node_feature_dim=IN_DIM, gat_out=D_OUT, heads=HEADS, num_segments=K_SEG, patch_size=PATCH,
                parallelism=f"shard-by-image x{n_gpus}",
                l2="working set per step (pool read + un-pool write) exceeds the 126 MB L2; no explicit flush")


# ---------------------------------------------------------------------------------------------
# CPU legs (cpu_baseline and --impl reference): the ONLY place bench.py touches oracle/
# ---------------------------------------------------------------------------------------------
class CpuArm:
    """The reference's per-image loop on the host cores.  kind "reference": the untouched reference classes
    (oracle/ref_block.RefGraphBlock over oracle/_ref) — Python double loop of construct_patch_graph, per-head loop,
    scatter_add_ and all; kind "port": oracle/restate.py (vectorised restatement), only when the copy is absent."""

    def __init__(self, name):
        import torch
        from oracle import ref_loader
        self.name = name
        self.H, self.W = WORKLOADS[name][:2]
        self.torch = torch
        if ref_loader.available():
            from oracle.ref_block import RefGraphBlock
            self.kind = "reference"
            self.where = os.path.relpath(ref_loader.REF_ROOT, ROOT) if ref_loader.REF_ROOT.startswith(ROOT) else ref_loader.REF_ROOT
            self.blk = RefGraphBlock(IN_DIM, 128, D_OUT, HEADS, K_SEG, PATCH, seed=1234)
            self.desc = (f"untouched reference classes from {self.where} driven as scripts/train_end_to_end.py:318-421 "
                         f"(oracle/ref_block.py), fp32, eval")
        else:
            from oracle import restate as O
            self.kind = "port"
            self.O = O
            self.params = O.init_block_params(IN_DIM, D_OUT, HEADS, K_SEG, seed=1234)
            self.desc = "oracle/restate.py (CPU restatement; the reference copy oracle/_ref is absent), fp32"

    def run(self, n_images, seed=0):
        """``n_images`` synthetic images of the workload through the per-image loop; returns seconds."""
        torch = self.torch
        gen = torch.Generator().manual_seed(seed)
        fms = [torch.randn(IN_DIM, self.H, self.W, generator=gen) for _ in range(n_images)]
        t0 = time.perf_counter()
        with torch.no_grad():
            for fm in fms:
                if self.kind == "reference":
                    self.blk.image(self.H, self.W, feature_map=fm, want_dense=True)
                else:
                    x = self.O.patch_mean_pool(fm, PATCH)
                    self.O.graph_block_image(x, self.H, self.W, self.params, K=K_SEG, want_dense=True)
        return time.perf_counter() - t0


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arm = CpuArm(args.workload)
    sample = args.cpu_sample_images or 2
    for _ in range(max(min(args.warmup, 3), 1)):
        arm.run(1)
    times = [arm.run(sample, seed=s) for s in range(args.steps)]
    total = sum(times)
    value = sample * args.steps / total
    cfg = workload_config(args.workload, args.gpus)
    line = {
        "impl": "reference", "metric": "graph_block_images_per_s", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": torch.get_num_threads(), "kind": arm.kind,
                         "sample": f"{sample} images of the workload per step; {arm.desc}"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """NVML sampling thread (SM clock + clock-event reasons) running during the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self, note):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0,
                    "note": "NVML unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "note": note}


class Ctx:
    """Per-process state shared by the legs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torch.distributed.run for N>1")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = "not bound (single rank: the cpu_baseline leg uses every host core)"
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            if args.numa_bind:
                # one process per GPU: keep this rank's pinned buffers and launch thread on the GPU's own NUMA node
                from mingraph_unet_b200.distributed import bind_to_gpu_numa
                self.numa = bind_to_gpu_numa(self.dev)
            else:
                self.numa = "not bound (--no-numa-bind)"
        # fail fast instead of hanging the box if an exchange never completes: a heartbeat the step functions advance;
        # the watchdog aborts the rank when it has not moved for 240 s (progress-based, so long --steps runs are fine)
        self.beat = [time.monotonic()]
        if self.world > 1:
            def _watch():
                while True:
                    time.sleep(5.0)
                    if time.monotonic() - self.beat[0] > 240.0:
                        sys.stderr.write("bench.py: rank %d made no progress for 240 s (stuck exchange?); aborting\n" % self.rank)
                        sys.stderr.flush()
                        os._exit(3)
            threading.Thread(target=_watch, daemon=True).start()

    def tick(self):
        self.beat[0] = time.monotonic()

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        import torch
        import torch.distributed as dist
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]


def make_block(dev, train=False):
    """The modules' own construction = the reference's init (xavier_uniform gain 1.414, graph_attention.py:36-37, same
    RNG consumption) under seed 1234."""
    import torch
    import mingraph_unet_b200 as mg
    torch.manual_seed(1234)
    blk = mg.GraphBlock(node_feature_dim=IN_DIM, gat_output_dim=D_OUT, num_heads=HEADS, num_segments=K_SEG).to(dev)
    return blk.train() if train else blk.eval()


# ---------------------------------------------------------------------------------------------
# forward leg (inference step): pool -> block (+ fused peer push) -> un-pool, pipelined CUDA-graph replays
# ---------------------------------------------------------------------------------------------
def forward_leg(ctx, name, B, steps, warmup, distributed=True, sampler=None, kernel_times=False, e2e=False):
    """Returns the measurements of one forward workload on this rank (times already max-reduced over ranks when
    ``distributed``).  ``distributed=False``: this rank alone, no exchange, no collective anywhere (the 1-GPU reference
    point of the strong-scaling efficiency)."""
    import torch
    import mingraph_unet_b200 as mg
    from mingraph_unet_b200 import _lib
    from mingraph_unet_b200.distributed import InlineGather, PeerExchange

    args, dev = ctx.args, ctx.dev
    world = ctx.world if distributed else 1
    H, W, _, dtname = WORKLOADS[name]
    dtype = getattr(torch, dtname)
    nph, npw = -(-H // PATCH), -(-W // PATCH)
    N = nph * npw
    E = 2 * (nph * (npw - 1) + npw * (nph - 1))
    depth = max(1, args.pipeline_depth)
    blk = make_block(dev)

    gen = torch.Generator().manual_seed(1000 + ctx.rank)
    fm_host = torch.randn(B, IN_DIM, H, W, generator=gen).to(dtype).pin_memory()
    fm_dev = fm_host.to(dev)
    # one fusion buffer per pipeline slot: [0:32] decoder features, [32:96] F_g
    fusions = [torch.zeros(B, C_UNET + D_OUT, H, W, dtype=dtype, device=dev) for _ in range(depth)]
    f_g_slices = [f[:, C_UNET:] for f in fusions]

    # public API: the block recorded once into a CUDA graph per pipeline slot (pool -> fused block kernel -> un-pool ->
    # [wait for every rank's payload]), replayed per step; `depth` independent steps in flight on round-robin streams so
    # the HBM-bound un-pool of step i overlaps the pool + latency-bound cluster kernel of step i+1.
    ex = ig = None
    mode = "none (1 GPU)"
    if world > 1:
        if args.exchange == "p2p":
            ex = PeerExchange(B, N, K_SEG, D_OUT, dev, depth)
            mode = "p2p: payload stored into every rank's gathered buffer by block_forward_kernel (NVLink peer memory), flag wait at the end of the step's graph; no collective"
        else:
            ig = InlineGather(B, N, K_SEG, D_OUT, dev, depth)
            mode = "inline: one NCCL all_gather_into_tensor per step from the step's stream"
    packed_small = ex.packed if ex is not None else ig.packed if ig is not None else [
        torch.zeros(B * (1 + K_SEG * D_OUT + N), dtype=torch.float32, device=dev) for _ in range(depth)]
    lc0 = _lib.launch_count()
    pipe = mg.PipelinedGraphBlock(blk, fm_dev, image_size=(H, W), outs=f_g_slices, shards=args.shards, depth=depth, warmup=2,
                                  packed_small=packed_small, peers=None if ex is None else ex.slots(),
                                  epilogues=None if ex is None else ex.epilogues())
    if ex is not None:
        ex.reset()                      # the graphs' warm-up passes pushed too: restart the step counters together
    runner = pipe.runners[0]
    per_step_kernels = (_lib.launch_count() - lc0 - 1) // (3 * depth)   # per slot: 2 warm-up passes + the recorded one (+1 weight prepare)

    def step(src=None):
        ctx.tick()
        slot, out = pipe.submit(src)
        if ig is not None:              # one collective enqueue from the step's own stream, no pack kernel
            ig.gather(slot, pipe.stream(slot))
            pipe.mark(slot)
        elif ex is not None:
            ex.stepped(slot)
        return slot, out

    def barrier():
        pipe.join()
        if distributed:
            ctx.barrier()
        else:
            torch.cuda.synchronize()

    # ---- device-resident timing (value) ------------------------------------------------------
    for _ in range(max(warmup, 3)):
        step()
    barrier()
    launches0 = _lib.launch_count()
    if sampler is not None:
        sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(steps):
        step()
    pipe.join()
    t_end.record()
    barrier()
    if sampler is not None:
        sampler.stop()
    ms_total = t_start.elapsed_time(t_end)
    res = {"B": B, "N": N, "E": E, "H": H, "W": W, "dtype": dtype, "depth": depth, "exchange": mode, "steps": steps,
           "per_step_kernels": per_step_kernels, "shards": runner.shards,
           "launches": (_lib.launch_count() - launches0) + per_step_kernels * steps}
    note = "sampled during the timed region"
    # very short timed region: keep the same load running while sampling (decided on the rank-reduced time: the steps
    # exchange, so every rank must run the same number of them)
    short = (ctx.max_over_ranks([ms_total])[0] if distributed else ms_total) < 150.0
    if sampler is not None and short:
        sampler.start()
        for _ in range(100):                                   # same count on every rank (the steps exchange)
            for _ in range(40):
                step()
            pipe.join()
            torch.cuda.synchronize()
        sampler.stop()
        note = "timed region shorter than the NVML sampling period; sampled under the same load right after it"
    res["clock_note"] = note

    # latency of ONE step (a single replay at a time), reported next to the pipelined throughput: rank-local graph
    # replays, so only where no exchange is recorded in the graph
    if ex is None:
        lat_steps = max(10, min(steps, 100))
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(lat_steps):
            runner()
        l1.record()
        torch.cuda.synchronize()
        res["latency_ms"] = l0.elapsed_time(l1) / lat_steps

    # ---- per-kernel durations: the same K steps launched eagerly with CUDA events around each kernel ----
    if kernel_times:
        kern_ev = {"pool": [], "block": [], "unpool": []}
        layers = (blk.patch_gat_model.gat_layers[0], blk.segment_predictor.gnn_predictor.gat_layers[0],
                  blk.region_gat_model.gat_layers[0])

        def ev():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e

        # launch shape = what the graph launches: one shard (B / shards images) per kernel
        Bs = runner._ranges[0][1] if runner.shards > 1 else B
        fm_shard, out_shard = fm_dev[:Bs], f_g_slices[0][:Bs]
        with torch.no_grad():
            prep = blk._prepared()
            for i in range(steps + 3):
                ctx.tick()
                torch.cuda._sleep(400_000)      # ~0.2 ms spin so the host runs ahead and the three launches queue back to back
                e0 = ev()
                x = mg.ops.pool_patches(fm_shard, PATCH, PATCH)
                e1 = ev()
                hh, SS, ll, lo, _, GG = mg.ops.block_forward(x, nph, npw, prep, D_OUT, layers[0].num_heads, layers[1].num_heads,
                                                             layers[2].num_heads, K_SEG)
                e2 = ev()
                mg.ops.unpool_nearest(GG, ll, nph, npw, H, W, out=out_shard)
                e3 = ev()
                if i >= 3:
                    kern_ev["pool"].append((e0, e1)); kern_ev["block"].append((e1, e2)); kern_ev["unpool"].append((e2, e3))
        torch.cuda.synchronize()
        res["kern_ms"] = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in kern_ev.items()}
        res["Bs"] = Bs

    # ---- end-to-end through the public API with host buffers (e2e) ---------------------------
    if e2e:
        host_loss = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(depth)]
        host_region = [torch.empty(B, K_SEG, D_OUT, dtype=torch.float32).pin_memory() for _ in range(depth)]
        host_labels = [torch.empty(B, N, dtype=torch.int32).pin_memory() for _ in range(depth)]

        def make_e2e(p, src):
            def e2e_step():
                ctx.tick()
                slot, out = p.submit(src)                               # H2D of this step's input from pinned memory
                if p is pipe:
                    if ig is not None:
                        ig.gather(slot, p.stream(slot))
                    elif ex is not None:
                        ex.stepped(slot)
                with torch.cuda.stream(p.stream(slot)):                 # D2H of the step's results, on the step's stream
                    host_loss[slot].copy_(out.l_partition, non_blocking=True)
                    host_region[slot].copy_(out.region_features, non_blocking=True)
                    host_labels[slot].copy_(out.hard_labels, non_blocking=True)
                p.mark(slot)
                # the host takes delivery of the OLDEST step in flight (the slot the next submit reuses): at most `depth`
                # steps are outstanding and every step's results are in host memory before the timed region closes
                p.host_wait((slot + 1) % depth)
            return e2e_step

        def time_e2e(p, fn):
            for _ in range(3):
                fn()
            p.join(); barrier()
            n = max(10, min(steps, 50))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            p.join()
            b.record()
            p.join(); barrier()
            return a.elapsed_time(b), n

        # (ii) the workload of `value`: full-resolution feature map in (what the conv decoder hands over), small outputs out
        res["e2e_ms"], res["e2e_steps"] = time_e2e(pipe, make_e2e(pipe, fm_host))
        res["e2e_h2d"] = fm_host.numel() * fm_host.element_size()
        res["e2e_d2h"] = 4 * (host_loss[0].numel() + host_region[0].numel() + host_labels[0].numel())
        # (i) the reference's literal block boundary (scripts/train_end_to_end.py:326): (B,N,in) node features in, the
        # same outputs out; no pooling, no exchange recorded (rank-local pipeline), same un-pool into the fusion buffers
        with torch.no_grad():
            x_dev = mg.ops.pool_patches(fm_dev, PATCH, PATCH)
        x_host = x_dev.cpu().pin_memory()
        pipe_nf = mg.PipelinedGraphBlock(blk, x_dev, image_size=(H, W), outs=f_g_slices, shards=1, depth=depth, warmup=2)
        res["e2e_nf_ms"], res["e2e_nf_steps"] = time_e2e(pipe_nf, make_e2e(pipe_nf, x_host))
        res["e2e_nf_h2d"] = x_host.numel() * x_host.element_size()
        del pipe_nf

    if distributed and ctx.world > 1:
        keys = ["ms_total"] + [k for k in ("e2e_ms", "e2e_nf_ms") if k in res]
        vals = [ms_total] + [res[k] for k in keys[1:]]
        if kernel_times:
            keys += ["k_pool", "k_block", "k_unpool"]
            vals += [res["kern_ms"]["pool"], res["kern_ms"]["block"], res["kern_ms"]["unpool"]]
        red = ctx.max_over_ranks(vals)
        ms_total = red[0]
        for k, v in zip(keys[1:], red[1:]):
            if k.startswith("k_"):
                res["kern_ms"][k[2:]] = v
            else:
                res[k] = v
    res["ms_total"] = ms_total
    res["ms_per_step"] = ms_total / steps
    res["images_per_s"] = B * world * steps / (ms_total * 1e-3)
    if ex is not None:
        res["exchange_status"] = int(ex.status.item())      # 0: no flag wait ran into its spin bound
        del pipe
        ex.close()
    return res


# ---------------------------------------------------------------------------------------------
# training leg (BASELINE configs[4]): forward + graph-layer backward scatter + NCCL gradient all-reduce + Adam
# ---------------------------------------------------------------------------------------------
def train_leg(ctx, B, steps, warmup, distributed=True, sampler=None, e2e=False):
    import torch
    import mingraph_unet_b200 as mg
    from mingraph_unet_b200 import _lib

    dev = ctx.dev
    world = ctx.world if distributed else 1
    H, W, _, dtname = WORKLOADS["cfg5"]
    dtype = getattr(torch, dtname)
    nph, npw = -(-H // PATCH), -(-W // PATCH)
    N = nph * npw
    E = 2 * (nph * (npw - 1) + npw * (nph - 1))
    blk = make_block(dev, train=True)          # dropout 0.1 on, as the reference trains (configs/model.yaml)
    opt = torch.optim.Adam(blk.parameters(), lr=1e-4, capturable=True, fused=True)   # one multi-tensor kernel (the foreach path is 56 launches)
    gen = torch.Generator().manual_seed(1000 + ctx.rank)
    fm_host = torch.randn(B, IN_DIM, H, W, generator=gen).to(dtype).pin_memory()
    fm_dev = fm_host.to(dev)
    # what the downstream heads send back: a dense cotangent d loss / d F_g (the conv stack stays stock PyTorch and is
    # outside the block), handed to the block's backward as is
    wdense = (torch.randn(B, D_OUT, H, W, generator=gen) / (H * W)).to(dtype).to(dev)
    host_loss = torch.empty(1, dtype=torch.float32).pin_memory()

    def loss_fn(out):                          # the block's own loss term; the dense map's gradient is the cotangent below
        return out.l_partition.mean()

    # public API: the whole step (pool -> block fwd -> loss -> backward from (dL/dF_g, loss) -> [NCCL all-reduce] -> Adam)
    # as CUDA graphs
    lc0 = _lib.launch_count()
    trainer = mg.CapturedTrainStep(blk, opt, fm_dev, (H, W), loss_fn, out_dtype=dtype, warmup=3, allreduce=distributed,
                                   dense_cotangent=wdense)
    trainer_launches = (_lib.launch_count() - lc0) // 4                         # 3 warm-up steps + 1 recorded step
    fm_in = trainer.static_in

    def step(src):
        ctx.tick()
        return trainer(None if src is fm_in else src)

    def barrier():
        if distributed:
            ctx.barrier()
        else:
            torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step(fm_in)
    barrier()
    l0 = _lib.launch_count()
    if sampler is not None:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step(fm_in)
    t1.record()
    barrier()
    if sampler is not None:
        sampler.stop()
    res = {"B": B, "N": N, "E": E, "H": H, "W": W, "dtype": dtype, "steps": steps, "trainer_launches": trainer_launches,
           "launches": (_lib.launch_count() - l0) + trainer_launches * steps}       # recorded kernels launch once per replay
    ms_total = t0.elapsed_time(t1)

    # the backward scatter kernel alone (dense gradient -> per-label rows), CUDA events on the launching stream
    labels = torch.randint(0, K_SEG, (B, N), device=dev, dtype=torch.int32)
    evs = []
    for i in range(23):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mg.ops.unpool_nearest_backward(wdense, labels, K_SEG, nph, npw)
        b.record()
        if i >= 3:
            evs.append((a, b))
    torch.cuda.synchronize()
    res["bwd_ms"] = statistics.mean(a.elapsed_time(b) for a, b in evs)

    if e2e:
        def e2e_step():
            loss = step(fm_host)
            host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(3):
            e2e_step()
        barrier()
        n = max(10, min(steps, 50))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            e2e_step()
        e1.record()
        barrier()
        res["e2e_ms"], res["e2e_steps"] = e0.elapsed_time(e1), n
        res["e2e_h2d"] = fm_host.numel() * fm_host.element_size()
    if distributed and ctx.world > 1:
        keys = ["bwd_ms"] + (["e2e_ms"] if e2e else [])
        red = ctx.max_over_ranks([ms_total] + [res[k] for k in keys])
        ms_total = red[0]
        for k, v in zip(keys, red[1:]):
            res[k] = v
    res["ms_total"] = ms_total
    res["ms_per_step"] = ms_total / steps
    res["images_per_s"] = B * world * steps / (ms_total * 1e-3)
    return res


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def solo_on_rank0(ctx, fn):
    """Run ``fn`` on rank 0 alone while the other ranks wait at a barrier; returns its result on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    out = None
    if ctx.rank == 0:
        out = fn()
        torch.cuda.synchronize()
    if ctx.world > 1:
        ctx.tick()
        dist.barrier()
    return out


CFG4_POINTS = ((262144, 8, 64), (262144, 32, 64), (262144, 8, 128), (262144, 32, 128))     # (nodes, in-degree, in = out features)


def gat_layer_points(dev, iters=8):
    """configs[3] points: one multi-head GAT layer (4 heads averaged, in = out = F, bf16 storage) on a random graph with fixed
    in-degree, timed with CUDA events on the launching stream after warm-up, L2 flushed (256 MB write) before every launch.
    Same generator and byte models as tools/sweep.py."""
    import torch
    from mingraph_unet_b200 import _lib, ops
    peak, _ = measured_peak()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    rows = []
    for N, k, F in CFG4_POINTS:
        gen = torch.Generator().manual_seed(0)
        tgt = torch.arange(N).repeat_interleave(k)
        src = torch.randint(0, N, (N * k,), generator=gen)
        rowptr, col, _ = ops.csr_from_coo(torch.stack([src, tgt]).to(dev), N, by_target=True)
        gen = torch.Generator().manual_seed(1)
        x = torch.randn(N, F, generator=gen).to(dev).to(torch.bfloat16)
        bound = 1.414 * (6.0 / (F + F)) ** 0.5
        W = ((torch.rand(4, F, F, generator=gen) * 2 - 1) * bound).to(dev)
        bound_a = 1.414 * (6.0 / (2 * F + 1)) ** 0.5
        a = ((torch.rand(4, 2 * F, generator=gen) * 2 - 1) * bound_a).to(dev)
        fn = lambda: ops.gat_forward(x, rowptr, col, W, a, concat=False, slope=0.2, out_dtype=torch.bfloat16)  # noqa: E731
        for _ in range(3):
            fn()
        launches0 = _lib.launch_count()
        ts = []
        for _ in range(iters):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        E = N * k
        comp = 2 * N * F * 2 + 4 * (E + N + 1) + 8 * 4 * N
        gath = E * F * 2 + N * F * 2 + 4 * (E + N + 1)
        rows.append({"N": N, "k": k, "F": F, "heads": 4, "E": E, "ms": ms, "edges_per_s": E / (ms * 1e-3),
                     "compulsory_gbs": comp / (ms * 1e-3) / 1e9, "compulsory_frac_of_hbm_peak": comp / (ms * 1e-3) / 1e9 / peak,
                     "gather_model_gbs": gath / (ms * 1e-3) / 1e9, "gather_model_frac_of_hbm_peak": gath / (ms * 1e-3) / 1e9 / peak,
                     "gpu_launches_per_layer": (_lib.launch_count() - launches0) // iters})
        del x, W, a, rowptr, col
    del flush
    torch.cuda.empty_cache()
    return rows


def extras(ctx, steps, warmup):
    """extra.cfg3 (configs[2], strong scaling), extra.cfg5 (configs[4], training step) and extra.cfg4 (configs[3] points) of the
    same invocation."""
    import gc
    import torch
    world = ctx.world
    out = {}
    steps3 = max(10, min(steps, 40))
    B3 = CFG3_GLOBAL_BATCH // world
    r = forward_leg(ctx, "cfg3", B3, steps3, 3)
    one = r if world == 1 else solo_on_rank0(ctx, lambda: forward_leg(ctx, "cfg3", CFG3_GLOBAL_BATCH, steps3, 3, distributed=False))
    if ctx.rank == 0:
        b = 2
        step_bytes = B3 * (D_OUT * r["H"] * r["W"] * b + IN_DIM * r["H"] * r["W"] * b)
        out["cfg3"] = {
            "workload": f"configs[2]: 1024x1024, global batch {CFG3_GLOBAL_BATCH} sharded by image, {B3} images/GPU x {world} GPUs, "
                        f"bf16 storage + fp32 math, forward graph block (same step as the headline)",
            "scaling": "strong", "images_per_s": r["images_per_s"], "ms_per_step": r["ms_per_step"], "steps": steps3,
            "images_per_s_1gpu_same_run": one["images_per_s"], "ms_per_step_1gpu_same_run": one["ms_per_step"],
            "efficiency": r["images_per_s"] / (world * one["images_per_s"]),
            "step_hbm_gbs_per_gpu": step_bytes / (r["ms_per_step"] * 1e-3) / 1e9,
            "exchange": r["exchange"], "exchange_status": r.get("exchange_status"),
            "gpu_launches": int(r["launches"]),
        }
    del r, one
    gc.collect(); torch.cuda.empty_cache()
    steps5 = max(10, min(steps, 100))
    B5 = WORKLOADS["cfg5"][2]
    t = train_leg(ctx, B5, steps5, 3)
    one = t if world == 1 else solo_on_rank0(ctx, lambda: train_leg(ctx, B5, steps5, 3, distributed=False))
    if ctx.rank == 0:
        peak, _ = measured_peak()
        bwd_bytes = B5 * (D_OUT * t["H"] * t["W"] * 2 + 4 * t["N"] + K_SEG * D_OUT * 4)
        out["cfg5"] = {
            "workload": f"configs[4]: TRAINING step at 512x512, {B5} images/GPU x {world} GPUs (batch 32 on 8 GPUs): pool -> block "
                        f"forward (dropout on) -> loss -> backward scatter -> NCCL all-reduce of the 22 792 graph-layer gradients -> Adam",
            "scaling": "weak", "images_per_s": t["images_per_s"], "ms_per_step": t["ms_per_step"], "steps": steps5,
            "images_per_s_1gpu_same_run": one["images_per_s"], "ms_per_step_1gpu_same_run": one["ms_per_step"],
            "efficiency": t["images_per_s"] / (world * one["images_per_s"]),
            "unpool_backward_ms": t["bwd_ms"], "unpool_backward_frac_of_hbm_peak": bwd_bytes / (t["bwd_ms"] * 1e-3) / 1e9 / peak,
            "gpu_launches": int(t["launches"]), "kernels_per_step": t["trainer_launches"],
        }
    gc.collect(); torch.cuda.empty_cache()
    pts = solo_on_rank0(ctx, lambda: gat_layer_points(ctx.dev))
    if ctx.rank == 0:
        out["cfg4"] = {
            "workload": "configs[3] points: one GAT layer (4 heads averaged, in = out = F, bf16 storage, random graph of fixed in-degree k) "
                        "on one GPU; median of 8 launches, L2 flushed before each; bytes as in tools/sweep.py",
            "hbm_peak_gbs": measured_peak()[0], "points": pts,
        }
    gc.collect(); torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    ctx = Ctx(args)
    world, rank = ctx.world, ctx.rank
    name = args.workload
    H, W, B, dtname = WORKLOADS[name]
    sampler = ClockSampler(ctx.local)
    if name == "cfg5":
        t = train_leg(ctx, B, args.steps, args.warmup, sampler=sampler, e2e=True)
        if rank == 0:
            peak, peak_src = measured_peak()
            cfg = workload_config(name, world)
            cfg["workload"] = cfg["workload"].replace("graph block (", "graph block TRAINING step (fwd + bwd + grad all-reduce + Adam; ")
            bwd_bytes = B * (D_OUT * H * W * 2 + 4 * t["N"] + K_SEG * D_OUT * 4)
            ach = bwd_bytes / (t["bwd_ms"] * 1e-3) / 1e9
            line = {
                "metric": "graph_block_train_images_per_s", "value": t["images_per_s"], "unit": "images/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 storage / f32 math",
                "data": "synthetic", "config": cfg, "edges_per_s": t["images_per_s"] * t["E"],
                "gpu_launches": int(t["launches"]),
                "launch_mode": "CUDA graph replay of the whole step (CapturedTrainStep); kernels of libmingraph_b200.so recorded "
                               "in the graph: %d per step" % t["trainer_launches"],
                "roofline": {"kernel": "un-pool backward scatter (dense gradient -> per-label rows)", "bound": "hbm",
                             "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": bwd_bytes, "kernel_ms": t["bwd_ms"]},
                "e2e": {"value": B * world * t["e2e_steps"] / (t["e2e_ms"] * 1e-3), "unit": "images/s",
                        "h2d_bytes_per_step": t["e2e_h2d"], "d2h_bytes_per_step": 4, "steps": t["e2e_steps"],
                        "api": "CapturedTrainStep(GraphBlock.train(), Adam): H2D + graph replay (+ NCCL all-reduce) + D2H of the loss"},
                "clocks": sampler.summary("sampled during the timed region"),
            }
            print(json.dumps(line), flush=True)
    else:
        r = forward_leg(ctx, name, B, args.steps, args.warmup, sampler=sampler, kernel_times=True, e2e=True)
        extra = {}
        if name == "cfg2" and not args.no_extras:
            extra = extras(ctx, args.steps, args.warmup)
        if rank == 0:
            peak, peak_src = measured_peak()
            cfg = workload_config(name, world)
            dtype, N, E, Bs = r["dtype"], r["N"], r["E"], r["Bs"]
            b = 2 if dtype == torch.bfloat16 else 4
            unpool_bytes = Bs * (D_OUT * H * W * b + 4 * N + K_SEG * D_OUT * 4)          # per launch (one shard)
            pool_bytes = Bs * (IN_DIM * H * W * b + N * IN_DIM * b)
            step_bytes = (unpool_bytes + pool_bytes) * B // Bs
            km = r["kern_ms"]
            achieved = unpool_bytes / (km["unpool"] * 1e-3) / 1e9
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                    traffic = json.load(f).get(name, {}).get("unpool_dram_bytes_per_launch_shards%d" % r["shards"])
            except Exception:
                pass
            line = {
                "metric": "graph_block_images_per_s", "value": r["images_per_s"],
                "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if dtype == torch.float32 else "bf16 storage / f32 math", "data": "synthetic", "config": cfg,
                "edges_per_s": r["images_per_s"] * E,
                "step_hbm_gbs": step_bytes / (r["ms_per_step"] * 1e-3) / 1e9,
                "gpu_launches": int(r["launches"]),
                "launch_mode": "CUDA graph replay (%d kernels of libmingraph_b200.so per step, %d parallel shard branches per "
                               "step, %d independent steps in flight on round-robin streams)" % (r["per_step_kernels"], r["shards"], r["depth"]),
                "pipeline_depth": r["depth"], "step_latency_ms": r.get("latency_ms"), "exchange": r["exchange"],
                "exchange_status": r.get("exchange_status"),
                "roofline": {"kernel": "unpool_vec_kernel (K7 nearest un-pool)", "bound": "hbm", "achieved": achieved,
                             "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": unpool_bytes,
                             "kernel_ms": km["unpool"], "frac_of_nominal_8TBs": achieved / 8000.0,
                             "images_per_launch": Bs,
                             "timing": "CUDA events around eager launches of the same kernels with the graph's launch shape "
                                       "(one shard of %d images; graph replays cannot carry timing events)" % Bs,
                             "other_kernels": {
                                 "pool_patches_tma_kernel": {"ms": km["pool"], "algorithmic_bytes": pool_bytes,
                                                             "achieved_gbs": pool_bytes / (km["pool"] * 1e-3) / 1e9,
                                                             "frac": pool_bytes / (km["pool"] * 1e-3) / 1e9 / peak},
                                 "block_forward_kernel": {"ms": km["block"], "note": "latency-bound cluster kernel; "
                                                          "moves ~%.1f MB" % (B * N * (IN_DIM * b + 4 * (D_OUT + 12)) / 1e6)}}},
                "e2e": {"value": B * world * r["e2e_steps"] / (r["e2e_ms"] * 1e-3), "unit": "images/s",
                        "h2d_bytes_per_step": r["e2e_h2d"], "d2h_bytes_per_step": r["e2e_d2h"],
                        "steps": r["e2e_steps"], "numa": ctx.numa,
                        "boundary": "feature map in: the (B,20,H,W) per-pixel map the conv decoder hands over crosses PCIe every step "
                                    "(the workload of `value`, pooling included)",
                        "api": "PipelinedGraphBlock(GraphBlock).submit(pinned host feature map): H2D + graph replay + D2H of "
                               "loss / region features / labels per step, %d steps in flight" % r["depth"]},
                "e2e_node_features": {"value": B * world * r["e2e_nf_steps"] / (r["e2e_nf_ms"] * 1e-3), "unit": "images/s",
                                      "h2d_bytes_per_step": r["e2e_nf_h2d"], "d2h_bytes_per_step": r["e2e_d2h"],
                                      "steps": r["e2e_nf_steps"],
                                      "boundary": "the reference's literal block input (scripts/train_end_to_end.py:326): (B,N,20) node "
                                                  "features in, same outputs out; no pooling kernel in the step, rank-local (no exchange)",
                                      "api": "PipelinedGraphBlock(GraphBlock).submit(pinned host node features)"},
                "clocks": sampler.summary(r["clock_note"]),
            }
            if extra:
                line["extra"] = extra
            if world == 1 and not args.no_cpu_baseline:
                cores = os.cpu_count() or 1
                torch.set_num_threads(cores)
                arm = CpuArm(name)
                arm.run(1)
                probe = arm.run(2) / 2                                              # s/image
                n_img = args.cpu_sample_images or int(min(4000, max(4, 12.0 / probe)))   # ~12 s of CPU work
                secs = arm.run(n_img)
                line["cpu_baseline"] = {"value": n_img / secs, "unit": "images/s", "cores": torch.get_num_threads(),
                                        "kind": arm.kind, "sample": f"{n_img} images of the workload in {secs:.1f} s; {arm.desc}"}
            print(json.dumps(line), flush=True)
    if world > 1:
        ctx.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
