#!/usr/bin/env python
"""Benchmark of the pgm-vae hot path on B200 (see BASELINE.md / SURVEY.md 8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg1]

A "step" is one training step (forward, MSE + VQ loss, backward, Adam, EMA codebook update) over one batch of
synthetic binary data.  Default workload = BASELINE.json configs[2] ("cfg3": 1556 variables, units 400/200/100/50,
K=512, D=64, EMA, batch 4096 per GPU -- the configuration the metric "train samples/s at 1/2/4/8 B200" is quoted
on; it fits one GPU: ~58 GB).  Prints ONE JSON line on rank 0.

  value       whole-job training samples/s, batches already resident in HBM (device pointers)
  e2e         the same through the public API (core.model.VqVAE.train_on_batch) with PINNED HOST batches: H2D copy
              of the batch and D2H read of the loss inside every step
  roofline    the dominant kernel of the step, from a per-kernel CUDA-event pass over the same steps (`traffic`: DRAM bytes
              of the launch named in `traffic_launch`, from the committed ncu capture profiles/r2_dram_traffic.json)
  dp_parity   N > 1 only, outside the timed region: three steps through the data-parallel path on every rank (wide
              models: the sharded peer-to-peer exchange fused with Adam) vs the same global batches on one GPU (rank 0)
              -- losses, weights, Adam moments, codebook, PLL counts, replicas bit-identical; the run FAILS above 1e-3
  cpu_baseline / --impl reference : the CPU restatement of the reference (oracle/, torch-CPU fp32; TensorFlow itself is
              not installable in this image) timed on the host cores on a bounded sample
  pll_eval    stage 2 (encoder + VQ assignment + histogram) samples/s: sample-sharded, variable-sharded, e2e
  cfg2        (N = 1) the round-1 headline workload (69 variables, chain kernels) as a secondary block
  vq_assign   BASELINE.json configs[3] (16 Mi vectors, D=64, K=8192): fused fp16 tcgen05 assignment + EMA scatter
  hbm_stages  the stand-alone EMA scatter / EMA update / PLL histogram kernels against the measured HBM peak
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "pgm-vae_b200"))

WORKLOADS = {
    # name: (V, units, D, K, per-GPU batch, description)
    "cfg1": (16, [15, 14, 13, 12], 4, 32, 256, "cfg1: nltcs shape, 16 vars, units 15/14/13/12, K=32 D=4 EMA, batch 256"),
    "cfg2": (69, [50, 40, 30, 20], 16, 128, 4096,
             "cfg2: synthetic 69-var binary data (plants-scale), units 50/40/30/20, K=128 D=16 EMA, batch 4096"),
    "cfg3": (1556, [400, 200, 100, 50], 64, 512, 4096,
             "cfg3: synthetic 1556-var binary data, units 400/200/100/50, K=512 D=64 EMA, batch 4096 per GPU"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def flops_per_sample(V, units, D, K):
    """dense fwd + bwd (no dgrad for layer 0) + VQ distance contraction."""
    chain = [V - 1] + units + [D] + units[::-1] + [V - 1]
    macs = [chain[i] * chain[i + 1] for i in range(10)]
    fwd = 2 * V * sum(macs)
    bwd = 2 * V * (2 * sum(macs) - macs[0])
    return fwd + bwd + 2 * V * D * K, 2 * V * (sum(macs[:5]) + D * K)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.path = None, f"/tmp/pgmvae_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [t.strip() for t in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- reference arm (CPU)
def cpu_train_rate(V, units, D, K, B, steps, warmup, budget_s, seed=0):
    """Times the oracle's training step (the CPU restatement of the reference) on the host cores.

    The V networks of the model are independent (one GEMM chain per variable), so for models whose full state does
    not fit a CPU step budget the sample is a SUBSET OF THE NETWORKS at the full layer widths: `nets` of the V nets,
    batch Bc, and the rate is scaled by nets / V (work is linear in the number of nets).  Returns
    (samples_per_s, seconds_per_sampled_step, cores, description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import pgmvae_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    train_fl, _ = flops_per_sample(V, units, D, K)

    def build(nets, Bc):
        rng = np.random.default_rng(seed)
        p = {}
        for i, (fin, fout) in enumerate(O.layer_dims(units, V, D)):
            lim = np.sqrt(6.0 / (V * fin)) if i < 9 else np.sqrt(6.0 / (V * fin + V * fout))
            p[f"fd{i}.kernel"] = torch.from_numpy(rng.uniform(-lim, lim, (nets, fin, fout)).astype(np.float32))
            p[f"fd{i}.bias"] = torch.zeros(nets, 1, fout)
        lim = np.sqrt(3.0 / (V * D))
        p["vq.embeddings"] = torch.from_numpy(rng.uniform(-lim, lim, (nets, D, K)).astype(np.float32))
        m = O.OracleVqVAE(units, nets, D, K, cost=0.25, decay=0.99, ema=True, params=p)
        y = O.synthetic_binary(2 * Bc, V, seed=seed)
        xs = [torch.from_numpy(np.stack([np.delete(y[i * Bc:(i + 1) * Bc], v, axis=1) for v in range(nets)], axis=1)
                               .astype(np.float32)) for i in range(2)]
        return m, xs

    # size the sample from a probe: ~flops the host does per second
    nets, Bc = min(V, 8), min(B, 64)
    m, xs = build(nets, Bc)
    m.train_step(xs[0], lr=1e-3)
    t0 = time.perf_counter()
    m.train_step(xs[1], lr=1e-3)
    t_probe = time.perf_counter() - t0
    rate = train_fl * Bc * nets / V / max(t_probe, 1e-6)                       # flop/s
    per_step = budget_s / max(steps + warmup, 1)
    want = rate * per_step                                                     # flops one sampled step may cost
    full = train_fl * B
    if want >= full:
        nets, Bc = V, B
    else:
        Bc = min(B, 256)
        nets = int(max(1, min(V, want / (train_fl * Bc / V))))
        if nets >= V:                                   # all networks fit: spend the rest of the budget on the batch
            Bc = int(min(B, max(Bc, want / train_fl)))
        if nets < 4:
            nets = min(V, 4)
            Bc = int(min(B, max(16, want / (train_fl * nets / V))))
    m, xs = build(nets, Bc)
    for i in range(warmup):
        m.train_step(xs[i % 2], lr=1e-3)
    t0 = time.perf_counter()
    for i in range(steps):
        m.train_step(xs[i % 2], lr=1e-3)
    dt = (time.perf_counter() - t0) / steps
    sps = Bc * (nets / V) / dt
    desc = (f"{steps} training steps of {nets} of the {V} per-variable networks (full layer widths) x batch {Bc} after {warmup} "
            f"warm-up (torch-CPU fp32 oracle, {cores} threads); samples/s scaled by {nets}/{V} (the networks are independent)")
    return sps, dt, cores, desc


def run_reference(args, wl):
    V, units, D, K, B, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sps, t_step, cores, sample = cpu_train_rate(V, units, D, K, B, args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": "train_samples_per_s", "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "per_gpu_batch": B,
                   "note": "CPU restatement of the reference (oracle/pgmvae_oracle.py, torch-CPU fp32); "
                           "TensorFlow, which the reference needs, is not installable in this image"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def committed(name):
    """a number taken from a committed ncu capture (profiles/<name>), or None"""
    p = os.path.join(ROOT, "profiles", name)
    return json.load(open(p)) if os.path.exists(p) else None


def vq_and_hbm_stages(ctx, _ffi, L, n_vq, hbm_peak, tc_peak_bf16):
    """cfg4-shaped VQ microbench (fused assign + EMA scatter, fp16 operands) and the HBM-bound stages alone."""
    import ctypes as C
    rng = np.random.default_rng(0)
    D, K = 64, 8192
    e = rng.uniform(-1, 1, (1, K, D)).astype(np.float32) * np.float32(np.sqrt(3.0 / D))
    z = rng.standard_normal((1, n_vq, D), dtype=np.float32)
    dz, de = _ffi.DeviceArray.from_numpy(ctx, z), _ffi.DeviceArray.from_numpy(ctx, e)
    del z
    idx = _ffi.DeviceArray(ctx, (1, n_vq), np.int32)
    cnt, dw = _ffi.DeviceArray(ctx, (1, K)), _ffi.DeviceArray(ctx, (1, K, D))

    def fused():
        _ffi.check(L.pgmvae_vq_assign_ema(ctx.h, None, dz.ptr, n_vq * D, D, de.ptr, K * D, D, idx.ptr, n_vq, cnt.ptr, K,
                                          dw.ptr, K * D, D, 1, n_vq, D, K))

    def plain():
        _ffi.check(L.pgmvae_vq_assign(ctx.h, None, dz.ptr, n_vq * D, D, de.ptr, K * D, D, idx.ptr, n_vq, None, None, 1, n_vq,
                                      D, K))
    out = {}
    ctx.set_precision(_ffi.PREC_BF16)
    for name, fn in (("fused_assign_ema", fused), ("assign_only", plain)):
        for _ in range(2):
            fn()
        ctx.sync()
        ctx.timer_start()
        reps = 3
        for _ in range(reps):
            fn()
        ms = ctx.timer_stop_ms() / reps
        tf = 2.0 * n_vq * D * K / (ms * 1e-3) / 1e12
        out[name] = {"ms": ms, "vectors_per_s": n_vq / (ms * 1e-3), "useful_tflops": tf, "frac_of_bf16_peak": tf / tc_peak_bf16}
    nres = C.c_int(0)
    _ffi.check(L.pgmvae_vq_assign_rescored(ctx.h, 1, K, C.byref(nres)))
    cap = committed("r2_vq_tensor_pipe.json")
    vq = {"workload": f"cfg4: {n_vq} vectors, D=64, K=8192, fp16 tcgen05 assignment (+ fused EMA scatter), exact indices",
          **out["fused_assign_ema"], "assign_only": out["assign_only"], "peak_tflops_bf16": tc_peak_bf16,
          "full_scan_rows": nres.value, "tensor_pipe_pct": cap,
          "note": "useful flops = 2*D*K per vector; the |e|^2 column adds 25 % MMA work that is not counted"}
    # stand-alone scatter + update on the same vectors / codes
    bc, bw = _ffi.DeviceArray(ctx, (1, K)), _ffi.DeviceArray(ctx, (1, K, D))
    ec, ew = _ffi.DeviceArray(ctx, (1, K)), _ffi.DeviceArray(ctx, (1, K, D))
    V, K2, B2 = 1556, 512, 32768
    idx2 = _ffi.DeviceArray.from_numpy(ctx, rng.integers(0, K2, (V, B2)).astype(np.int32))
    y2 = _ffi.DeviceArray.from_numpy(ctx, (rng.random((B2, V)) < 0.2).astype(np.uint8))
    n1, n0 = _ffi.DeviceArray(ctx, (V, K2), np.uint64), _ffi.DeviceArray(ctx, (V, K2), np.uint64)

    def stages():
        _ffi.check(L.pgmvae_ema_stats(ctx.h, None, dz.ptr, n_vq * D, D, idx.ptr, n_vq, cnt.ptr, K, dw.ptr, K * D, D, 1,
                                      n_vq, D, K))
        _ffi.check(L.pgmvae_ema_apply(ctx.h, None, cnt.ptr, dw.ptr, bc.ptr, bw.ptr, ec.ptr, ew.ptr, de.ptr, 1, K, D, D,
                                      0.99, 1e-5, 1, 1))
        _ffi.check(L.pgmvae_pll_count(ctx.h, None, idx2.ptr, B2, y2.ptr, V, 0, n1.ptr, n0.ptr, V, B2, K2))
    stages()
    ctx.sync()
    ctx.profile_begin()
    for _ in range(3):
        stages()
    hbm = {}
    for k in ctx.profile_end():
        gbs = k["bytes"] / max(k["ms"], 1e-9) / 1e6
        hbm[k["name"]] = {"ms_per_launch": k["ms"] / k["launches"], "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
    hbm["shapes"] = {"ema": f"{n_vq} x 64 vectors into 8192 codes", "pll_count": f"{V} variables x {B2} samples, K={K2}"}
    return vq, hbm


# --------------------------------------------------------------------------- our arm
def time_training(args, ctx, L, _ffi, model, y_dev, B, gB, V, nb, comm_h, barrier, max_over_ranks, min_ms=400.0):
    """W warm-up steps, then K * R timed steps (R repeats of the --steps loop so that the timed region is at least
    ~0.4 s: a 12 ms sample is noise at 8 GPUs); device timed, max over ranks."""
    import ctypes as C
    lr = C.c_float(1e-3)

    def step_dev(i, met=None):
        off = (i % nb) * B * V
        _ffi.check(L.pgmvae_model_train_step(model._h, y_dev.ptr + off, 1, B, gB, lr, comm_h, 0, met))
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    ctx.timer_start()
    for i in range(args.steps):
        step_dev(i)
    probe = max_over_ranks(ctx.timer_stop_ms())
    R = int(max(1, min(200, np.ceil(min_ms / max(probe, 1e-3)))))
    barrier()
    l0 = ctx.launches
    ctx.timer_start()
    for i in range(args.steps * R):
        step_dev(i)
    ms = ctx.timer_stop_ms()
    launches = ctx.launches - l0
    barrier()
    ms = max_over_ranks(ms)
    return ms, R, launches, step_dev


def run_ours(args, wl, wl_name):
    V, units, D, K, B, desc = wl
    from pgmvae import _ffi, data, dist
    from core.model import VqVAE, Adam
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = _ffi.get_context(local)
    prec = {"tf32": _ffi.PREC_TF32, "bf16": _ffi.PREC_BF16, "fp32": _ffi.PREC_FP32}[args.precision]
    ctx.set_precision(prec)
    comm, rank, world = dist.init_from_env(ctx)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    L = _ffi.lib()

    def barrier():
        ctx.sync()
        if world > 1:
            import torch.distributed as tdist
            tdist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch
        import torch.distributed as tdist
        t = torch.tensor([x], dtype=torch.float64)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t[0])

    # ---- data-parallel parity of exactly the path that is timed below (outside the timed region)
    dp_par = None
    if world > 1 and not args.no_dp_parity:
        sys.path.insert(0, os.path.join(ROOT, "pgm-vae_b200", "tools"))
        import dp_check
        dp_par = dp_check.dp_parity(ctx, comm, rank, world, units, V, D, K, min(B, 256), args.precision, ema=True, local=local,
                                    eval_samples=4096)
        ctx.set_precision(prec)
        if rank == 0:
            dp_par["failed"] = dp_check.check(dp_par)

    model = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=True, seed=0, max_batch=B, device=local, comm=comm)
    model.compile(optimizer=Adam(lr=1e-3), loss="mse", metrics=["mae"])
    arith = {0: "fp32", 1: "tf32", 2: "bf16"}[int(L.pgmvae_model_arithmetic(model._h))]
    gB = B * world
    nb = 8                                              # distinct batches, rotated
    y = data.synthetic_binary(nb * B, V, seed=1000 + rank)
    y_dev = _ffi.DeviceArray.from_numpy(ctx, y)
    comm_h = comm.h if comm is not None else None

    # ---- device-resident timing ("value")
    sampler = ClockSampler(local) if rank == 0 else None
    ms, R, launches, step_dev = time_training(args, ctx, L, _ffi, model, y_dev, B, gB, V, nb, comm_h, barrier, max_over_ranks)
    clocks = sampler.stop() if sampler else None
    nsteps = args.steps * R
    value = gB * nsteps / (ms * 1e-3)
    met = (C.c_double * 4)()
    step_dev(0, met)
    assert all(np.isfinite(v) for v in met), "non-finite loss"

    # ---- end-to-end through the public API, pinned host batches
    hp = C.c_void_p()
    _ffi.check(L.pgmvae_malloc_host(ctx.h, nb * B * V, C.byref(hp)))
    pinned = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nb * B, V))
    pinned[:] = y
    for i in range(max(3, args.warmup)):
        model.train_on_batch(pinned[(i % nb) * B:(i % nb + 1) * B], global_batch=gB, sync=True)
    barrier()
    ctx.timer_start()
    t0 = time.perf_counter()
    for i in range(nsteps):
        model.train_on_batch(pinned[(i % nb) * B:(i % nb + 1) * B], global_batch=gB, sync=True)
    e2e_ms = ctx.timer_stop_ms()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(max(e2e_ms, e2e_wall))
    e2e_value = gB * nsteps / (e2e_ms * 1e-3)

    # ---- per-kernel pass (CUDA events around every launch of the same steps) -> roofline of the top kernel
    ctx.profile_begin()
    psteps = min(args.steps, 10)
    for i in range(psteps):
        step_dev(i)
    prof = ctx.profile_end()
    hbm_peak, tc_peak_bf16, peak_src = peaks()
    tot_ms = sum(k["ms"] for k in prof) or 1.0
    prof.sort(key=lambda k: -k["ms"])
    top = prof[0]
    avg_ms = top["ms"] / top["launches"]
    gbs = top["bytes"] / top["launches"] / (avg_ms * 1e-3) / 1e9
    tfs = top["flops"] / top["launches"] / (avg_ms * 1e-3) / 1e12
    intensity = top["flops"] / max(top["bytes"], 1.0)
    # the measured tensor peak is bf16; tf32 runs at half that rate
    tc_peak = tc_peak_bf16 / 2 if arith == "tf32" else tc_peak_bf16
    tensor_bound = intensity > (tc_peak * 1e12) / (hbm_peak * 1e9) and ("_tc" in top["name"] or "_bf16" in top["name"])
    traffic = committed("r2_dram_traffic.json")
    roofline = {
        "kernel": top["name"], "bound": "tensor" if tensor_bound else "hbm",
        "achieved": tfs if tensor_bound else gbs, "peak": tc_peak if tensor_bound else hbm_peak,
        "unit": "TFLOP/s" if tensor_bound else "GB/s",
        "frac": (tfs / tc_peak) if tensor_bound else (gbs / hbm_peak),
        "traffic": (traffic or {}).get(wl_name, {}).get(top["name"]) if traffic else None,
        "traffic_source": "ncu --set full capture committed as profiles/r2_dram_traffic.json (per launch)" if traffic else None,
        "traffic_launch": (traffic or {}).get("_launch", {}).get(top["name"]) if traffic else None,
        "peak_source": peak_src + (" bf16 sustained" if arith != "tf32" else " bf16 sustained / 2 (tf32 operands)"),
        "avg_launch_ms": avg_ms, "launches_per_step": top["launches"] / psteps,
        "share_of_step": top["ms"] / tot_ms, "algorithmic_bytes_per_launch": top["bytes"] / top["launches"],
        "algorithmic_flops_per_launch": top["flops"] / top["launches"],
        "method": "CUDA events around every library launch over %d of the timed steps (separate pass)" % psteps,
        "kernels": [{"name": k["name"], "ms_per_step": k["ms"] / psteps, "share": k["ms"] / tot_ms,
                     "GBps": k["bytes"] / max(k["ms"], 1e-9) / 1e6, "TFLOPs": k["flops"] / max(k["ms"], 1e-9) / 1e9}
                    for k in prof],
    }

    # ---- stage 2: PLL evaluation (encoder + assignment + histogram); cfg5 = the cfg3 model over 10 M samples, here a
    # bounded sample of it per GPU
    train_fl, pll_fl = flops_per_sample(V, units, D, K)
    n_eval = nb * B
    n1 = np.zeros((V, K), np.uint64)
    n0 = np.zeros((V, K), np.uint64)
    reps = 2
    _ffi.check(L.pgmvae_model_count(model._h, y_dev.ptr, 1, n_eval, n1.ctypes.data, n0.ctypes.data))
    barrier()
    ctx.timer_start()
    for _ in range(reps):
        _ffi.check(L.pgmvae_model_count(model._h, y_dev.ptr, 1, n_eval, n1.ctypes.data, n0.ctypes.data))
    pll_ms = max_over_ranks(ctx.timer_stop_ms())
    assert (n1 + n0).sum() == n_eval * V
    # variable-sharded: every rank evaluates its own V / world networks on the same n_eval samples
    v0, v1 = model.var_shard()
    barrier()
    ctx.timer_start()
    for _ in range(reps):
        _ffi.check(L.pgmvae_model_count_vars(model._h, y_dev.ptr, 1, n_eval, v0, v1, n1.ctypes.data, n0.ctypes.data))
    pllv_ms = max_over_ranks(ctx.timer_stop_ms())
    t0 = time.perf_counter()
    model.dist = model.cpt(pinned)
    pll = model.pseudo_log_likelihood(pinned, total=n_eval * world)
    pll_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / 2.0)      # two passes over the data
    pll_value = world * n_eval * reps / (pll_ms * 1e-3)
    pll_eval = {"metric": "pll_eval_samples_per_s", "value": pll_value, "unit": "samples/s",
                "variable_sharded_value": n_eval * reps / (pllv_ms * 1e-3),
                "e2e_value": world * n_eval / (pll_e2e_ms * 1e-3),
                "samples_per_pass": world * n_eval, "pll": pll,
                "tensor_roofline_frac": pll_value * pll_fl / 1e12 / (world * tc_peak),
                "note": ("cfg5 (BASELINE.json configs[4]) is this model over 10 M samples; this is a bounded sample of it, "
                         "%d samples per GPU and pass" % n_eval) if wl_name == "cfg3" else None}
    _ffi.check(L.pgmvae_free_host(ctx.h, hp))

    if rank != 0:
        return
    device_bytes = model.device_bytes()
    groups = -(-V // model.group_size())
    del model, y_dev
    vq_assign = hbm_stages = cfg2 = None
    if world == 1 and not args.no_microbench:
        vq_assign, hbm_stages = vq_and_hbm_stages(ctx, _ffi, L, args.vq_n, hbm_peak, tc_peak_bf16)
    if world == 1 and wl_name != "cfg2" and not args.no_secondary:
        # the round-1 headline workload (chain kernels, tf32) beside it
        V2, u2, D2, K2, B2, desc2 = WORKLOADS["cfg2"]
        ctx.set_precision(_ffi.PREC_TF32)
        m2 = VqVAE(u2, V2, D2, K2, cost=0.25, decay=0.99, ema=True, seed=0, max_batch=B2, device=local)
        m2.compile(optimizer=Adam(lr=1e-3))
        y2 = data.synthetic_binary(nb * B2, V2, seed=1000)
        y2d = _ffi.DeviceArray.from_numpy(ctx, y2)
        a2 = argparse.Namespace(steps=50, warmup=5)
        ms2, R2, _, _ = time_training(a2, ctx, L, _ffi, m2, y2d, B2, B2, V2, nb, None, barrier, max_over_ranks)
        fl2, _ = flops_per_sample(V2, u2, D2, K2)
        v2 = B2 * 50 * R2 / (ms2 * 1e-3)
        cfg2 = {"workload": desc2, "value": v2, "unit": "samples/s", "ms_per_step": ms2 / (50 * R2), "steps": 50 * R2,
                "dtype": "tf32", "achieved_tflops": v2 * fl2 / 1e12}
        del m2, y2d
        ctx.set_precision(prec)
    # ---- CPU baseline beside it (rank 0, N == 1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sps, t_step, cores, sample = cpu_train_rate(V, units, D, K, B, 3, 1, budget_s=20.0)
        cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample}
    step_tf = value * train_fl / 1e12
    line = {
        "metric": "train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "repeats": R, "timed_steps": nsteps, "warmup": args.warmup, "ms_per_step": ms / nsteps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": arith, "data": "synthetic",
        "config": {"workload": desc, "per_gpu_batch": B, "precision": arith + " operands, fp32 accumulate",
                   "global_batch": gB, "parallelism": f"dp{world}", "variable_groups_per_step": groups,
                   "l2": "per-step working set (activations + gradients: GBs) exceeds the 126 MB L2; 8 distinct input "
                         "batches rotated",
                   "flop_per_sample_train": train_fl, "flop_per_sample_pll": pll_fl,
                   "achieved_tflops": step_tf, "step_tensor_roofline_frac": step_tf / (world * tc_peak),
                   "device_bytes": device_bytes},
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": B * V, "d2h_bytes_per_step": 32,
                "ms_per_step": e2e_ms / nsteps},
        "gpu_launches": launches, "clocks": clocks, "pll_eval": pll_eval, "dp_parity": dp_par, "cfg2": cfg2,
        "vq_assign": vq_assign, "hbm_stages": hbm_stages,
        "loss_after": {"loss": met[0], "mse": met[1], "mae": met[2], "vq_loss": met[3]},
    }
    emit(line)
    if dp_par and dp_par.get("failed"):
        print(f"dp_parity FAILED: {dp_par['failed']}", file=sys.stderr)
        sys.exit(3)


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout, but libraries write there too (NCCL prints its version banner to
    stdout when the communicator comes up).  From here on file descriptor 1 is stderr; the result line goes to the
    original stdout through emit()."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    text = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(text.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, text)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-microbench", action="store_true", help="skip the cfg4 VQ and HBM-stage micro-benchmarks")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg2 block")
    ap.add_argument("--no-dp-parity", action="store_true", help="skip the data-parallel parity check (N > 1)")
    ap.add_argument("--vq-n", type=int, default=1 << 24, help="vectors of the cfg4 VQ micro-benchmark (16 Mi)")
    ap.add_argument("--precision", default="bf16", choices=["fp32", "tf32", "bf16"],
                    help="arithmetic of the GEMM-shaped kernels: bf16 = the fastest tensor-core path for the geometry (bf16 "
                         "tcgen05 for wide networks, tf32 chain kernels for narrow ones); tf32; or exact fp32 CUDA cores")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl, args.workload)


if __name__ == "__main__":
    main()
