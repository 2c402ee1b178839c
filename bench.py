#!/usr/bin/env python
"""bench.py — throughput of the ensemble-SAE training hot path (BASELINE.json metric: activations/sec/GPU).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own CPU PyTorch path

A "step" is one ``FunctionalEnsemble.step_batch`` over one batch of synthetic activations: forward, losses,
backward and the Adam update of every model of the ensemble (nothing is skipped or cached). The workload at N=1 is
BASELINE config 2: 16 tied SAEs, d_model=512, dict_ratio=8 (n=4096), L1 = logspace(-4,-2,16), batch 8192, fp32
parameters, lr 1e-3. For N>1 every rank trains its own 16-model shard on the same activation stream (config 4:
model-axis sharding, no data-path collective) — weak scaling; value = rows consumed by all ranks' shards per second.

Printed JSON (one line, rank 0): the driver contract plus
  roofline      dominant kernel (weight-gradient GEMM): algorithmic FLOPs / CUDA-event time vs the measured bf16 peak
  cpu_baseline  the oracle port of the reference step on this box's host cores, bounded sample
  e2e           same metric through the public API with HOST (pinned) batches: H2D copy + step + D2H of the losses
  phases_ms     per-phase device time of a step (events recorded inside libsce on the launching stream)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (M, d, n, B, description)
    "cfg2": (16, 512, 4096, 8192, "16 TiedSAE d_model=512 dict_ratio=8 L1=logspace(-4,-2,16) batch=8192 (BASELINE configs[1])"),
    "cfg1": (1, 128, 256, 1024, "1 TiedSAE d_model=128 dict_ratio=2 L1=1e-3 batch=1024 (BASELINE configs[0])"),
    "cfg5": (1, 2048, 32768, 4096, "1 TiedSAE/GPU d_model=2048 dict_ratio=16 batch=4096 (BASELINE configs[4])"),
    "cfg3": (12, 768, 6144, 8192, "12 TopK d_model=768 dict_ratio=8 k in {16,32,64} batch=8192 (one shape group of BASELINE configs[2])"),
}
METRIC = "activations/sec (whole job; rows consumed by every resident model)"


def l1_grid(M):
    return [1e-3] if M == 1 else [float(a) for a in np.logspace(-4, -2, M)]


def make_models(sig, M, d, n, seed):
    torch.manual_seed(seed)
    if getattr(sig, "variant", None) == "topk":
        return [sig.init(d, n, (16, 32, 64)[i % 3]) for i in range(M)]
    return [sig.init(d, n, a) for a in l1_grid(M)]


ACT_FP16 = True   # --act-precision: values as the reference caches them (fp16, activation_dataset.py:404-412) or raw fp32


def synth_batches(n_batches, B, d, seed, pin=False):
    """Sparse-mixture activations (the distribution of sc_datasets/random_dataset.py:76-142): a few unit features
    per row + noise, so that ReLU sparsity is non-trivial. Returns CPU fp32 tensors; with ACT_FP16 the VALUES are
    rounded to fp16 first, which is what a chunk written by the reference's harvester and loaded by big_sweep.py
    (`torch.load(chunk_loc).to(device="cpu", dtype=torch.float32)`, big_sweep.py:358) contains."""
    gen = torch.Generator().manual_seed(seed)
    feats = torch.randn(2048, d, generator=gen)
    feats /= feats.norm(dim=-1, keepdim=True)
    out = []
    for _ in range(n_batches):
        codes = (torch.rand(B, 2048, generator=gen) < 0.01).float() * torch.rand(B, 2048, generator=gen)
        x = codes @ feats + 0.05 * torch.randn(B, d, generator=gen)
        if ACT_FP16:
            x = x.half().float()
        out.append(x.pin_memory() if pin else x)
    return out


class ClockSampler:
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    @staticmethod
    def _ts(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t_begin=None, t_end=None):
        """Samples are kept only if their timestamp lies inside [t_begin, t_end] (the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            ts = self._ts(f[0])
            if t_begin is not None and ts is not None and not (t_begin - 0.02 <= ts <= t_end + 0.02):
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


class _StdoutGuard:
    """stdout must carry exactly ONE JSON line. Libraries (NCCL's version banner, for one) write to file descriptor
    1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON goes to a saved duplicate of
    the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, obj):
        sys.stdout.flush()
        os.write(self.real, (json.dumps(obj) + "\n").encode())


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "hbm_gbs": p.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    return json.load(open(path)) if os.path.exists(path) else None


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own PyTorch path (oracle port) on host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(M, d, n, B_full, budget_s=20.0, steps=1, warmup=1):
    """Times ``steps`` steps of the restated reference (vmap(grad(loss)) + Adam, fp32, all host threads) on a
    bounded sample: the full ensemble at a reduced batch chosen so that the work fits the budget. Returns
    (activations/s, description, cores)."""
    from oracle import sae_oracle as O
    from sparse_coding_b200 import FunctionalTiedSAE
    ncpu = os.cpu_count() or 1
    models = make_models(FunctionalTiedSAE, M, d, n, 0)
    ens = O.RefPortEnsemble(models, O.SIG_LOSSES["tied"], lr=1e-3)
    # probe at a small batch: pick the thread count that serves the reference best (oversubscribing SMT siblings
    # can be slower than fewer threads), then size the sample from its per-row time
    Bp = min(B_full, 256)
    chunk = synth_batches(1, max(Bp, 64), d, 123)[0]
    idx = torch.randperm(chunk.shape[0])[:Bp]
    torch.set_num_threads(ncpu)
    ens.step_batch(chunk[idx])                       # one-off tracing / allocator warm-up, not timed
    best = None
    for th in sorted({ncpu, max(1, ncpu // 2), min(ncpu, 32), min(ncpu, 16)}, reverse=True):
        torch.set_num_threads(th)
        t0 = time.perf_counter()
        ens.step_batch(chunk[idx])
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, th)
    probe, cores = best
    torch.set_num_threads(cores)
    per_row = probe / Bp
    Bs = int(min(B_full, max(Bp, budget_s / max(steps + warmup, 1) / per_row)))
    Bs = max(64, (Bs // 64) * 64)
    # bound the [M,B,n] fp32 temporaries (about 12 live copies) to ~24 GB of host memory
    Bs = min(Bs, max(64, int(24e9 / (12 * 4 * M * n)) // 64 * 64))
    chunk = synth_batches(1, Bs, d, 124)[0]
    for _ in range(warmup):
        ens.step_batch(chunk[torch.randperm(Bs)])      # includes the reference's CPU gather (big_sweep.py:168)
    t0 = time.perf_counter()
    for _ in range(steps):
        ens.step_batch(chunk[torch.randperm(Bs)])
    dt = (time.perf_counter() - t0) / steps
    return Bs / dt, f"{steps} step(s) of the full {M}-model ensemble at batch {Bs} of {B_full} rows (fp32, torch CPU)", cores, dt


def run_reference(args, rank, world, out):
    if rank != 0:
        return
    M, d, n, B, desc = WORKLOADS[args.workload]
    rate, sample, cores, dt = cpu_reference_rate(M, d, n, B, budget_s=60.0, steps=max(args.steps, 1),
                                                 warmup=max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "activations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "parallelism": "host CPU threads"},
        "cpu_baseline": {"value": rate, "unit": "activations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "activations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    out.emit(line)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--bwd-passes", type=int, default=3, choices=[1, 3])
    ap.add_argument("--arith", default="auto", choices=["auto", "bf16x3", "f16f8"],
                    help="operand arithmetic (include/sce.h sce_arith); auto = f16f8 where the shape allows")
    ap.add_argument("--act-precision", default="fp16", choices=["fp16", "fp32"],
                    help="synthetic activation VALUES: fp16-representable (the reference's chunk format; default) or "
                         "arbitrary fp32. The tensors fed to the engine are fp32 either way.")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the informational single-pass-backward run")
    args = ap.parse_args()
    global ACT_FP16
    ACT_FP16 = args.act_precision == "fp16"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    out = _StdoutGuard()
    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch.distributed as dist
    import sparse_coding_b200 as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout; stdout must carry exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    M, d, n, B, desc = WORKLOADS[args.workload]
    K, W = args.steps, max(args.warmup, 3)

    # every rank owns its own shard of the sweep: same shapes, different seeds (model-axis sharding)
    sig = S.TopKEncoder if args.workload == "cfg3" else S.FunctionalTiedSAE
    ens = S.FunctionalEnsemble(make_models(sig, M, d, n, seed=rank), sig, S.adam,
                               {"lr": 1e-3}, device=dev, bwd_passes=args.bwd_passes, arith=args.arith)
    n_pool = 8
    host = synth_batches(n_pool, B, d, seed=1000, pin=True)        # identical stream on every rank
    pool = [x.to(dev) for x in host]                                 # resident copies for the device-timed run

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident run: `value`
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(W):
        ens.step_batch(pool[i % n_pool])
    barrier()
    ens.profile_begin()
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e0.record()
    for i in range(K):
        losses, aux = ens.step_batch(pool[i % n_pool])
        launches += ens.gpu_launches_last_call()
    e1.record()
    barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    phases = ens.profile_end()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    final_loss = losses["loss"].detach().clone()

    # ---------------- end-to-end run through the public API with host batches: `e2e`
    barrier()
    h2d = B * d * 4
    d2h = 0
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(2):
        ens.step_batch(host[i % n_pool])
    barrier()
    e2.record()
    for i in range(K):
        losses, aux = ens.step_batch(host[i % n_pool])               # pinned host -> device copy inside
        got = {k: v.cpu() for k, v in losses.items()}                 # D2H of the step's result
        nnz = aux["c"].count_nonzero(dim=-1).float().mean(dim=-1).cpu()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    d2h = sum(v.numel() * 4 for v in got.values()) + nnz.numel() * 4

    # ---------------- informational: the same workload with single-pass bf16 backward GEMMs (NOT the headline)
    ms_alt = float("nan")
    if world == 1 and args.bwd_passes == 3 and sig is S.FunctionalTiedSAE and not args.no_alt:
        del pool[4:]
        alt = S.FunctionalEnsemble(make_models(S.FunctionalTiedSAE, M, d, n, seed=rank), S.FunctionalTiedSAE, S.adam,
                                   {"lr": 1e-3}, device=dev, bwd_passes=1, arith=args.arith)
        for i in range(3):
            alt.step_batch(pool[i % len(pool)])
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e4.record()
        for i in range(K):
            alt.step_batch(pool[i % len(pool)])
        e5.record()
        barrier()
        ms_alt = e4.elapsed_time(e5)
        del alt

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # end-of-chunk metric gather (the only collective on this path): every model's final loss to every rank
        gathered = [torch.empty_like(final_loss) for _ in range(world)]
        dist.all_gather(gathered, final_loss)
        final_loss = torch.cat(gathered)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        pk = peaks()
        arith = ens.resolved_arith()
        # tensor work issued per fp32-equivalent GEMM, in bf16-pass equivalents: three kind::f16 passes, or one
        # kind::f16 pass + two kind::f8f6f4 passes at twice the rate
        full_passes = 3 if arith == "bf16x3" else 2
        bwd_eq = full_passes if args.bwd_passes == 3 else 1
        # f16f8 + fp16-representable activations: x has no residual plane, so the x.l8 * W.h8 term of encode and the
        # dz.h8 * x.l8 term of the dz^T x half of dW are skipped on the device (one 8-bit pass = 1/2 pass equivalent)
        x_skip = arith == "f16f8" and ACT_FP16
        enc_eq = full_passes - (0.5 if x_skip else 0.0)
        dw_eq = (bwd_eq - (0.25 if x_skip else 0.0)) if args.bwd_passes == 3 else 1
        arith_text = {
            "bf16x3": "fp32 parameters/moments/accumulation; every GEMM operand is an exact-to-2^-17 (hi, lo) bf16 pair "
                      "and every product 3 tensor-core passes (hi*hi + hi*lo + lo*hi)",
            "f16f8": "fp32 parameters/moments/accumulation; every GEMM operand is an fp16 plane plus two e5m2 planes "
                     "(value, scaled residual); every product = one kind::f16 pass (h*h) + two kind::f8f6f4 passes for "
                     "the cross terms (2 bf16-pass equivalents), rescaled in the accumulator",
        }[arith] + "; parity <= 1e-4 rel vs the fp32 reference on x_hat and losses"
        value = world * B * K / (ms * 1e-3)
        e2e_value = world * B * K / (ms_e2e * 1e-3)
        steps_prof = max(phases["steps"], 1)
        per_phase = {k: v / steps_prof for k, v in phases.items() if k != "steps"}
        dw_ms = per_phase["dw"]
        alg_flops_dw = 4.0 * M * B * n * d            # dW = dz^T x + c^T g: two GEMMs of 2*B*n*d per model
        achieved = alg_flops_dw / (dw_ms * 1e-3) / 1e12 if dw_ms > 0 else None
        step_flops = 10.0 * M * B * n * d
        line = {
            "metric": METRIC, "value": value, "unit": "activations/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic (sparse mixture + noise); values " +
                    ("rounded to fp16 as the reference caches activations (activation_dataset.py:404-412), "
                     if ACT_FP16 else "arbitrary fp32, ") + "fed as fp32 tensors",
            "config": {"workload": f"{args.workload}: {desc}", "models_per_gpu": M,
                       "activation_values": args.act_precision, "d_model": d, "dict_size": n,
                       "batch": B, "parallelism": f"ensemble-shard x{world}" if world > 1 else "single GPU",
                       "arith": arith, "arithmetic": arith_text, "pass_equivalents_per_gemm": full_passes,
                       "x_residual_term_skipped": bool(x_skip),
                       "fwd_passes": 3, "bwd_passes": args.bwd_passes, "adam_count_mode": "frozen_t1",
                       "l2": "per-step working set (code + code-gradient, 4.3 GB) and the 8-batch input pool "
                             "(134 MB) both exceed the 126 MB L2; no explicit flush"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "activations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K},
            "roofline": {"bound": "tensor", "kernel": "gemm_split_kernel<EpiStoreF32,MN,MN> (weight gradient)",
                         "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / pk["bf16_tflops"] if achieved else None, "traffic": None,
                         "peak_source": pk["source"], "alg_flops_per_launch": alg_flops_dw,
                         "ms_per_launch": dw_ms,
                         "issued_tflops": alg_flops_dw * dw_eq / (dw_ms * 1e-3) / 1e12 if dw_ms > 0 else None,
                         "issued_note": f"{dw_eq} bf16-pass equivalents per fp32 FLOP of this kernel ({arith}"
                                        + (", x residual term skipped" if x_skip else "") + f"): frac <= 1/{dw_eq} x "
                                        "(tensor-pipe utilisation = issued_tflops / peak; `peak` is cuBLAS's sustained "
                                        "bf16 rate under the power cap, which kind::f8f6f4 passes can exceed)",
                         "step_alg_tflops": step_flops / (ms / K * 1e-3) / 1e12},
            "phases_ms": per_phase,
            # every GEMM phase against the same peak: algorithmic (fp32-equivalent) and issued (x passes) TFLOP/s
            "gemms": {ph: {"alg_tflops": units * 2.0 * M * B * n * d / (per_phase[ph] * 1e-3) / 1e12,
                           "issued_tflops": units * 2.0 * M * B * n * d * passes / (per_phase[ph] * 1e-3) / 1e12,
                           "frac_of_peak_issued": units * 2.0 * M * B * n * d * passes / (per_phase[ph] * 1e-3) / 1e12 / pk["bf16_tflops"]}
                      for ph, units, passes in (("encode", 1, enc_eq), ("decode", 1, full_passes), ("dcode", 1, bwd_eq),
                                                ("dw", 2, dw_eq)) if per_phase[ph] > 0},
            "final_loss_mean": float(final_loss.mean()),
        }
        tr = (ncu_traffic() or {}).get(arith)
        if tr:
            line["roofline"]["traffic"] = tr["dw_dram_bytes_per_launch"]
            line["roofline"]["traffic_source"] = tr["source"]
            # dz and c at 4 B / element (3 B for dz when x's residual term is skipped: its h8 plane is not read) + dW
            line["roofline"]["alg_bytes_per_launch"] = (8.0 - (1.0 if x_skip else 0.0)) * M * B * n + 4.0 * M * n * d
        if ms_alt == ms_alt:
            line["alt_precision"] = {"note": "informational only: backward GEMMs on the 16-bit plane alone (bwd_passes=1); "
                                             "forward, losses and x̂ unchanged",
                                     "value": B * K / (ms_alt * 1e-3), "ms_per_step": ms_alt / K}
        if world == 1 and not args.no_cpu_baseline:
            rate, sample, cores, _ = cpu_reference_rate(M, d, n, B, budget_s=20.0)
            line["cpu_baseline"] = {"value": rate, "unit": "activations/s", "cores": cores, "kind": "port",
                                    "sample": sample}
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
