#!/usr/bin/env python
"""bench.py -- document graphs/s of the fused CAGGC+MAGGC graph blocks (fwd+bwd) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass (forward + backward, all gradients) of the graph hot path over one shard of
synthetic DocRED-shaped document graphs that is resident in HBM: BASELINE.json configs[1]
(GCGCN_glove graph blocks, L_s=2, H=8, fp32) with the 12-document batch of SURVEY.md section 8d
(n = 42,35,...,5) tiled --tile times per GPU so that one step streams far more than the 126 MB L2.
With N > 1 every rank owns its own shard (documents are independent: weak scaling, no data-path
collective) and the ranks all-reduce the parameter-gradient bucket once per step (training use).

Prints ONE JSON line (rank 0).  `value` = documents of all ranks / max-over-ranks device time.
`e2e` = the same step driven from pinned HOST buffers: inputs copied host->device and results
device->host inside the timed region.  `roofline` describes the dominant kernel (per-kernel CUDA
events recorded by the library on the launching stream during the timed region) and the whole
path (`path_*`: SURVEY section 8d algorithmic bytes per document / step time).  `cpu_baseline` is
the oracle (CPU restatement of the reference, pinned bit-exact to it) timed on this box's cores.

--impl reference times that same CPU oracle as its own arm (the reference is PyTorch code that
cannot travel to the GPU box; oracle/gcgcn_oracle.py issues the same ATen ops).
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "document graphs/sec (fwd+bwd graph blocks)"
UNIT = "graphs/s"
VARIANTS = {"glove": (2, 8), "bert": (4, 4)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gcgcn_b200", choices=["gcgcn_b200", "reference"])
    ap.add_argument("--variant", default="glove", choices=list(VARIANTS))
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"], help="edge-tensor storage")
    ap.add_argument("--tile", type=int, default=512, help="12-document batches per GPU per step")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------- CPU oracle arm
def cpu_oracle_rate(variant: str, seconds: float, max_passes: int = 1000, warmup: int = 1):
    """graphs/s of the CPU oracle, documents one at a time like the reference trainer (C:339),
    fwd+bwd, all host threads."""
    import torch
    from helpers import blocks_state, oracle_blocks
    from gcgcn_b200 import synthetic

    layers, heads = VARIANTS[variant]
    torch.set_num_threads(os.cpu_count() or 1)
    _, state = blocks_state(layers, heads)
    docs = synthetic.make_batch()
    for _ in range(warmup):
        for d in docs:
            oracle_blocks(d, state, layers, heads)
    times = []
    t_end = time.perf_counter() + seconds
    while len(times) < max_passes and (time.perf_counter() < t_end or len(times) < 2):
        t0 = time.perf_counter()
        for d in docs:
            oracle_blocks(d, state, layers, heads)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return len(docs) / med, {"cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "passes": len(times),
                             "sample": f"{len(times)} passes over the 12-document batch (SURVEY 8d), "
                                       f"fwd+bwd, one document at a time, median pass {med * 1e3:.1f} ms"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    layers, heads = VARIANTS[args.variant]
    import torch
    from helpers import blocks_state, oracle_blocks
    from gcgcn_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    _, state = blocks_state(layers, heads)
    docs = synthetic.make_batch()
    for _ in range(max(args.warmup, 1)):
        for d in docs:
            oracle_blocks(d, state, layers, heads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for d in docs:
            oracle_blocks(d, state, layers, heads)
    dt = time.perf_counter() - t0
    value = args.steps * len(docs) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "variant": args.variant, "layer_num": layers,
                   "head_num": heads, "step": "reference arm: one step = one pass over the 12-document "
                                             "batch on the host cores (bounded sample of the workload)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "12-document batch per step, fwd+bwd, one document at a time; "
                                   "oracle/gcgcn_oracle.py (same ATen ops as the reference, pinned bit-exact)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(args):
    return (f"configs[1]: GCGCN_{args.variant} graph blocks fwd+bwd, 12-document DocRED-shaped batch "
            f"(n=42..5, SURVEY 8d) x {args.tile} = {12 * args.tile} documents per GPU per step")


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.tmp.read().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        self.tmp.close()
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------- roofline helpers
def kernel_bytes(name: str, bt, H: int, esz: int) -> float:
    """Minimum HBM bytes one launch of a kernel must move (its operands once), for the batch."""
    n1, n2, d = bt.total_nodes, bt.total_pairs, 128
    hd = H * d
    table = {
        "edge_row_fwd<score+mean>": n2 * d * esz + n2 * 4 + n1 * d * 4,           # read e0; write A, ebar
        "edge_row_fwd<mean>": n2 * d * esz + n1 * d * 4,                           # read e1; write ebar
        "edge_row_bwd<score+mean>": 2 * n2 * d * esz + n2 * 4 + n1 * d * 4,        # read e0, dS, debar; write de0
        "edge_row_bwd<mean>": n2 * d * esz + n1 * d * 4,                           # write de1
    }
    return float(table.get(name, 0.0))


# ------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from gcgcn_b200 import _lib, synthetic
    from gcgcn_b200.batch import RaggedBatch
    from gcgcn_b200.modules import GraphBlocks
    from gcgcn_b200.sharding import GradBucket

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gcgcn_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL's version banner goes to stdout at NCCL_DEBUG=VERSION
        os.environ["NCCL_DEBUG"] = os.environ.get("GCGCN_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    layers, heads = VARIANTS[args.variant]
    edt = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    esz = 4 if args.dtype == "fp32" else 2
    torch.manual_seed(0)
    gb = GraphBlocks(layers, heads).to(dev).eval()
    sizes = synthetic.shard_doc_sizes(12 * args.tile)
    bt = RaggedBatch(sizes, dev)
    ndocs = bt.num_docs

    # synthetic inputs of SURVEY 8d's distribution, generated on the device (seeded per rank)
    gen = torch.Generator(device=dev).manual_seed(1337 + rank)
    x0 = torch.tanh(torch.randn(bt.total_nodes, 128, device=dev, generator=gen)).requires_grad_(True)
    e0 = torch.randn(bt.total_pairs, 128, device=dev, generator=gen).to(edt).requires_grad_(True)
    e1 = torch.randn(bt.total_pairs, 128, device=dev, generator=gen).to(edt).requires_grad_(True)
    dy1 = torch.randn(bt.total_nodes, 128, device=dev, generator=gen)
    dy2 = torch.randn(bt.total_nodes, 128, device=dev, generator=gen)
    params = [p for n, p in gb.named_parameters() if "linears_k" not in n]
    bucket = GradBucket(params)

    def step():
        x0.grad = e0.grad = e1.grad = None
        for p in params:
            p.grad = None
        out = gb(x0, e0, e1, bt)
        torch.autograd.backward([out["y1"], out["y2"]], [dy1, dy2])
        if world > 1:
            bucket.pack()
            bucket.all_reduce()
            bucket.unpack()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, with_kernel_events=False):
        barrier()
        stream = torch.cuda.current_stream().cuda_stream
        launches0 = _lib.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if with_kernel_events:
            _lib.timing_begin(stream)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        kern = _lib.timing_end(stream) if with_kernel_events else {}
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, _lib.launch_count() - launches0, kern

    for _ in range(max(args.warmup, 3)):
        step()
    # per-kernel breakdown: an eager pass with the library's per-launch events (not the headline)
    ms_eager, launches, kern = timed(step, args.steps, with_kernel_events=True)
    graphed = None
    if not args.no_graph:
        from gcgcn_b200.graphs import GraphedPass
        x0.grad = e0.grad = e1.grad = None
        for p in params:
            p.grad = None

        def fwd_bwd():
            out = gb(x0, e0, e1, bt)
            torch.autograd.backward([out["y1"], out["y2"]], [dy1, dy2])
            return out

        graphed = GraphedPass(fwd_bwd, dev)

        def step():                                   # noqa: F811  (the timed step from here on)
            out = graphed.replay()
            if world > 1:
                bucket.pack()
                bucket.all_reduce()
                bucket.unpack()
            return out

        for _ in range(max(args.warmup, 3)):
            step()
    sampler = ClockSampler(local) if rank == 0 else None
    ms, _, _ = timed(step, args.steps)
    clocks = sampler.stop() if sampler else {}
    value = world * ndocs * args.steps / (ms * 1e-3)

    # ---- end to end from pinned host buffers (same step; H2D of inputs and D2H of results inside)
    e2e = None
    if not args.no_e2e:
        host_in = [t.detach().cpu().pin_memory() for t in (x0, e0, e1, dy1, dy2)]
        dev_in = [x0, e0, e1, dy1, dy2]
        host_out = {k: torch.empty(bt.total_nodes, 128).pin_memory() for k in ("y1", "y2", "dx0")}
        host_grads = torch.empty(bucket.flat.numel()).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in host_in)
        d2h = 3 * bt.total_nodes * 128 * 4 + host_grads.numel() * 4

        def e2e_step():
            with torch.no_grad():
                for h, dv in zip(host_in, dev_in):
                    dv.copy_(h, non_blocking=True)
            out = step()
            if world == 1:
                bucket.pack()
            host_out["y1"].copy_(out["y1"].detach(), non_blocking=True)
            host_out["y2"].copy_(out["y2"].detach(), non_blocking=True)
            host_out["dx0"].copy_(x0.grad, non_blocking=True)
            host_grads.copy_(bucket.flat, non_blocking=True)

        e2e_step()
        e2e_steps = max(2, min(args.steps, 5))
        ms_e, _, _ = timed(e2e_step, e2e_steps)
        e2e = {"value": world * ndocs * e2e_steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e / e2e_steps,
               "note": "inputs x0,e0,e1,dy1,dy2 from pinned host memory; y1,y2,dx0 and the parameter-gradient "
                       "bucket read back; de0/de1 stay on the device (their consumer, the edge-feature "
                       "producer's backward, lives there)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = read_peaks()
    step_s = ms * 1e-3 / args.steps
    alg = bt.algorithmic_bytes(esz, backward=True)
    path_gbs = alg / step_s / 1e9
    top = None
    total_k = sum(v[1] for v in kern.values()) or 1.0
    ranked = sorted(((k, v) for k, v in kern.items() if not k.startswith("(")), key=lambda kv: -kv[1][1])
    breakdown = [{"kernel": k, "launches": v[0], "ms_per_step": v[1] / args.steps, "share": v[1] / total_k}
                 for k, v in ranked[:30]]
    gaps = kern.get("(between calls)", (0, 0.0))
    breakdown.append({"kernel": "(time between C-ABI calls: torch packing/autograd glue)", "launches": gaps[0],
                      "ms_per_step": gaps[1] / args.steps, "share": gaps[1] / total_k})
    if ranked:
        name, (cnt, tot_ms) = ranked[0]
        per_launch_s = tot_ms * 1e-3 / cnt
        kb = kernel_bytes(name, bt, heads, esz)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(name)
        top = {"bound": "hbm", "kernel": name, "achieved": kb / per_launch_s / 1e9 if kb else 0.0, "peak": peak,
               "unit": "GB/s", "frac": (kb / per_launch_s / 1e9 / peak) if kb else 0.0, "traffic": traffic,
               "bytes_per_launch": kb, "us_per_launch": per_launch_s * 1e6, "share_of_step": tot_ms / total_k,
               "peak_source": peak_src,
               "path_bytes_per_step": alg, "path_achieved": path_gbs, "path_frac": path_gbs / peak,
               "path_note": "whole step: SURVEY 8d algorithmic bytes s*d*(5n^2+8n) per document / step time",
               "kernels": breakdown}

    cpu = None
    if not args.no_cpu_baseline:
        rate, info = cpu_oracle_rate(args.variant, args.cpu_seconds)
        cpu = {"value": rate, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"],
               "host_cpus": info["host_cpus"]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "variant": args.variant, "layer_num": layers, "head_num": heads,
                   "edge_storage": args.dtype, "docs_per_gpu": ndocs, "total_nodes": bt.total_nodes,
                   "total_pairs": bt.total_pairs, "parallelism": f"doc-sharded dp{world}",
                   "l2": f"inputs larger than L2: {2 * bt.total_pairs * 128 * esz / 1e9:.2f} GB of edge features "
                         "streamed per step",
                   "collective": "none" if world == 1 else f"one NCCL all-reduce of {bucket.nbytes} B per step",
                   "launch": "eager (one C-ABI call per op)" if graphed is None else
                             "CUDA-graph replay of the captured forward+backward pass (same kernels as eager)",
                   "eager_ms_per_step": ms_eager / args.steps},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": top, "cpu_baseline": cpu,
    }
    sys.stdout.flush()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
