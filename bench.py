#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native AttnLRP hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port on the host cores)

metric  : attributions/sec at seq 2048 on Llama-3-8B dims (BASELINE.json), synthetic random-init bf16 weights and
          random token ids.  One "step" = one pass of the whole path (embed -> forward -> arg-max logit -> LRP
          backward -> Gradient x Input) over `--per-gpu-batch` prompts per GPU.
value   : whole-job attributions/s, token ids already resident in HBM, device-timed (CUDA events, max over ranks).
e2e     : the same through the public API `LlamaAttnLRPEngine.attribute(host_ids) -> host relevance`
          (pinned-host ids copied H2D and relevance copied D2H inside the timed region).
roofline: dominant kernel = the tcgen05 GEMM (every Linear fwd / LRP dgrad); achieved = sum(2*M*N*K) / sum(CUDA-event
          duration) over all GEMM launches of the timed region, against the measured cuBLAS peak.
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
sys.path.insert(0, os.path.join(ROOT, "lrp-explains-transformers_b200"))

import torch  # noqa: E402

METRIC = "attributions/sec (seq2048) Llama-3-8B"
UNIT = "attributions/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=32)
    ap.add_argument("--micro-batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--model", default="llama3-8b", choices=["llama3-8b", "tinyllama-1.1b", "gemma3-4b", "llama-test"])
    ap.add_argument("--layers", type=int, default=0, help="override the number of layers (debug only; invalidates the metric)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def model_dims(name, layers=0):
    from lxt_b200.engine import GEMMA3_4B, LLAMA3_8B, TINYLLAMA_1B, LlamaDims, gemma3_dims
    import dataclasses
    if name == "gemma3-4b":   # non-headline workload (BASELINE configs[4]): use with --seq 8192 --per-gpu-batch 4 --micro-batch 1
        return gemma3_dims(d=2560, I=10240, H=8, Hkv=4, D=256, L=layers or 34, V=262208)
    d = {"llama3-8b": LLAMA3_8B, "tinyllama-1.1b": TINYLLAMA_1B,
         "llama-test": LlamaDims(d=512, I=1024, H=8, Hkv=2, D=64, L=2, V=1024)}[name]
    if layers:
        d = dataclasses.replace(d, L=layers)
    return d


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, smax, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(dims, seq, threads=None):
    """Time the oracle port (oracle/attnlrp_oracle.py, the CPU restatement of lxt.efficient + HF Llama) on the host
    cores on a bounded sample: ONE decoder layer at the full model width, batch 1, bf16, and extrapolate to L layers.
    Returns (attributions_per_s, cores, sample_description, seconds_per_layer)."""
    from oracle import attnlrp_oracle as O
    cores = threads or best_thread_count()
    torch.set_num_threads(cores)
    cfg = dict(d=dims.d, I=dims.I, H=dims.H, Hkv=dims.Hkv, D=dims.D, L=1, V=2048, eps=dims.eps, theta=dims.theta)
    w = O.random_llama_weights(cfg, seed=0)
    ids = torch.randint(0, cfg["V"], (1, seq), generator=torch.Generator().manual_seed(1))
    O.llama_attnlrp(w, ids, cfg, dtype=torch.bfloat16)  # warm-up
    t0 = time.perf_counter()
    O.llama_attnlrp(w, ids, cfg, dtype=torch.bfloat16)
    dt = time.perf_counter() - t0
    per_attr = dt * dims.L
    desc = (f"oracle port (torch CPU bf16, {cores} threads): 1 of {dims.L} decoder layers at full width "
            f"(d={dims.d}, I={dims.I}, H={dims.H}/{dims.Hkv}), S={seq}, B=1, fwd + LRP bwd = {dt:.2f} s; x{dims.L} layers")
    return 1.0 / per_attr, cores, desc, dt


def best_thread_count():
    """torch's CPU bf16 GEMM does not always scale to every hardware thread: time one [2048,4096]x[4096,4096] bf16
    matmul at a few thread counts and keep the fastest, so that the CPU baseline is not handicapped."""
    n = os.cpu_count() or 1
    cands = sorted({c for c in (n, n // 2, n // 4, 32, 16, 8) if 1 <= c <= n}, reverse=True)
    a = torch.randn(2048, 4096).bfloat16()
    b = torch.randn(4096, 4096).bfloat16()
    best, best_t = n, float("inf")
    for c in cands:
        torch.set_num_threads(c)
        a @ b
        t0 = time.perf_counter()
        for _ in range(3):
            a @ b
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dims = model_dims(args.model, args.layers)
    vals = []
    for i in range(args.warmup + args.steps):
        v, cores, desc, dt = cpu_reference_sample(dims, args.seq)
        if i >= args.warmup:
            vals.append(v)
        if dt * (args.warmup + args.steps) > 240:  # keep the whole arm within a few minutes
            vals = vals or [v]
            break
    v = sum(vals) / len(vals)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.model} random-init bf16, seq {args.seq}, CPU sample", "seq_len": args.seq},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ our arm
def run_b200(args):
    from lxt_b200 import dist as ldist, ops
    from lxt_b200.engine import LlamaAttnLRPEngine

    rank, world, local = ldist.init_from_env("nccl")
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dims = model_dims(args.model, args.layers)
    S, Bg = args.seq, args.per_gpu_batch
    eng = LlamaAttnLRPEngine.random_init(dims, device=dev, seed=0, micro_batch=args.micro_batch)
    N_total = Bg * world
    ids_all = torch.randint(0, dims.V, (N_total, S), generator=torch.Generator().manual_seed(1))
    lo, hi = ldist.shard_range(N_total, rank, world)
    ids_host = ids_all[lo:hi].contiguous().pin_memory()
    ids_dev = ids_host.to(dev)
    rel_host = torch.empty((hi - lo, S), dtype=torch.float32, pin_memory=True)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_device():
        rels = [eng.attribute_device(ids_dev[i:i + eng.micro_batch]) for i in range(0, hi - lo, eng.micro_batch)]
        rel = torch.cat(rels, 0)
        return ldist.gather_relevance(rel, N_total, world)  # the single NCCL collective of the path

    def step_e2e():
        out = eng.attribute(ids_host, out=rel_host)  # public API: host ids -> host relevance
        if world > 1:
            ldist.gather_relevance(out.to(dev, non_blocking=True), N_total, world)
        return out

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: device resident
    ops.GEMM_PROFILE = []
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    launches = ops.launch_count() - l0
    prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
    ms_dev = e0.elapsed_time(e1)
    gemm_flops = sum(p[0] for p in prof)
    gemm_ms = sum(p[1].elapsed_time(p[2]) for p in prof)
    # ---- timed region 2: end to end through the public API (host buffers)
    step_e2e()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
    if rank != 0:
        return
    value = N_total * args.steps / (ms_dev / 1e3)
    e2e_v = N_total * args.steps / (ms_e2e / 1e3)

    peaks, peak_src = None, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
            peak_src = "measured"
    except Exception:
        pass
    peak_tf = (peaks or {}).get("bf16_tflops_sustained") or 1400.0
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    metric = METRIC if (args.model == "llama3-8b" and S == 2048 and not args.layers) else f"attributions/sec (seq{S}) {args.model} [non-headline workload]"
    out = {
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.model} random-init bf16, seq {S}, {Bg} prompts per GPU per step "
                               f"(micro-batch {args.micro_batch}), batch-sharded over {world} GPU(s), 1 NCCL gather",
                   "global_batch": N_total, "seq_len": S, "layers": dims.L, "parallelism": f"dp{world}",
                   "l2": "inputs larger than L2 (16 GB weights + activation store streamed every step)"},
        "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(ids_host.numel() * 8),
                "d2h_bytes_per_step": int(rel_host.numel() * 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "gemm_bf16_pair_kernel (tcgen05 cta_group::2, all Linear fwd + LRP dgrad)",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                     "launches": len(prof), "share_of_step": gemm_ms / ms_dev if ms_dev else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of the largest-share launch shape
                     # (M=16384 N=28672 K=4096, gate|up forward) from profiles/r01_ncu_summary.md: 1.45 GB + 0.92 GB
                     # per launch vs 1.31 GB algorithmic (A 134 MB + W 235 MB + C 940 MB): W panels re-streamed 4x via L2
                     "traffic": 2.37e9, "traffic_note": "bytes/launch, ncu --set full, gate|up fwd shape; algorithmic 1.31e9"},
    }
    if args.model == "llama3-8b":
        try:
            out["kernels"] = secondary_kernel_rooflines(dims, S, dev, peaks)
        except Exception as ex:  # pragma: no cover
            out["kernels"] = {"error": str(ex)}
    if not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N=1 only
        try:
            v, cores, desc, _ = cpu_reference_sample(dims, S)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        except Exception as ex:  # pragma: no cover
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(out))


def secondary_kernel_rooflines(dims, S, dev, peaks):
    """The two kernels BASELINE.json's metric text names besides attributions/s, timed alone (CUDA events, after warm-up,
    outside the timed region): the fused relevance-space eps-LRP Linear rule (one launch, 4*T*K*N flops) against the measured
    burst bf16 peak, and flash AttnLRP forward/backward at the step's shape against both the HBM and the tensor roofline."""
    from lxt_b200 import ops
    burst = (peaks or {}).get("bf16_tflops") or 1590.0
    hbm = (peaks or {}).get("hbm_gbs") or 6650.0
    src = "measured" if peaks else "fallback"

    def timed(fn, n=5):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    res = {}
    T = 8 * S
    g = torch.Generator(device=dev).manual_seed(5)
    eps_rows = []
    for (K, N) in ((dims.d, dims.d), (dims.d, dims.I)):
        x = (torch.rand(T, K, generator=g, device=dev) + 0.5).to(torch.bfloat16)
        w = (torch.rand(N, K, generator=g, device=dev) * 0.02).to(torch.bfloat16)
        r = torch.randn(T, N, generator=g, device=dev)
        ms = timed(lambda: ops.linear_eps_bwd(x, w, None, r, 1e-6))
        tf = 4.0 * T * K * N / (ms * 1e-3) / 1e12
        eps_rows.append({"T": T, "K": K, "N": N, "ms": ms, "achieved": tf, "frac": tf / burst})
        del x, w, r
    res["eps_linear_fused"] = {"bound": "tensor", "unit": "TFLOP/s", "flops": "4*T*K*N (z = xW^T, s = R/(z+eps), s W, * x in ONE launch)",
                               "peak": burst, "peak_source": f"{src} bf16_tflops (burst: kernel timed alone)", "shapes": eps_rows,
                               "note": "time includes the workspace zero-fill + bf16 casts of the Python wrapper"}
    B, H, Hkv, D = 8, dims.H, dims.Hkv, dims.D
    qkv = torch.randn(B, S, (H + 2 * Hkv) * D, generator=g, device=dev).to(torch.bfloat16)
    q, k, v = qkv[:, :, : H * D].view(B, S, H, D), qkv[:, :, H * D: (H + Hkv) * D].view(B, S, Hkv, D), qkv[:, :, (H + Hkv) * D:].view(B, S, Hkv, D)
    d_o = torch.randn(B, S, H, D, generator=g, device=dev).to(torch.bfloat16)
    o, lse = ops.attn_fwd(q, k, v, D ** -0.5)
    gq = torch.empty_like(qkv)
    dq, dk, dv = gq[:, :, : H * D].view(B, S, H, D), gq[:, :, H * D: (H + Hkv) * D].view(B, S, Hkv, D), gq[:, :, (H + Hkv) * D:].view(B, S, Hkv, D)
    acc, dl = torch.empty(B, S, H, D, device=dev), torch.empty(B, H, S, device=dev)
    ms_f = timed(lambda: ops.attn_fwd(q, k, v, D ** -0.5))
    ms_b = timed(lambda: ops.attn_bwd(q, k, v, o, d_o, lse, D ** -0.5, dq=dq, dk=dk, dv=dv, dq_acc=acc, delta=dl))
    fl_f = 4.0 * B * H * S * S * D / 2
    # read Q, K, V, O, dO, LSE; write dQ, dK, dV (bf16; LSE fp32)  -> 84.2 MB per Llama-3-8B layer-sequence
    bytes_fb = B * (4 * S * H * D * 2 + 4 * S * Hkv * D * 2 + S * H * 4)
    res["flash_attnlrp"] = {"shape": {"B": B, "S": S, "H": H, "Hkv": Hkv, "D": D, "causal": True},
                            "fwd_ms": ms_f, "bwd_ms": ms_b, "fwd_tflops": fl_f / (ms_f * 1e-3) / 1e12,
                            "bwd_tflops": 2.5 * fl_f / (ms_b * 1e-3) / 1e12, "tensor_peak": burst,
                            "algorithmic_bytes_fwd_bwd": bytes_fb, "hbm_gbs_achieved": bytes_fb / ((ms_f + ms_b) * 1e-3) / 1e9,
                            "hbm_peak_gbs": hbm, "hbm_frac": bytes_fb / ((ms_f + ms_b) * 1e-3) / 1e9 / hbm,
                            "note": "at S=2048 the kernel is arithmetic/latency bound (~1000 FLOP/B), not HBM bound"}
    return res


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
