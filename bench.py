#!/usr/bin/env python
"""bench.py -- CD-k training throughput of the iDBN hot loop on B200 (BASELINE.json configs[1], "C2").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision tf32x2|tf32|fp32]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" = one minibatch of ``iDBN.train`` on [10000,1500,500], CD-1, batch 64 per GPU (reference
idbn.py:199-204): for each layer one ``train_epoch`` and one ``forward`` with the updated weights.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.

Default precision: 'tf32x2' -- the tcgen05 kernels with split operands, which meet the SAME parity bars as fp32
(tests/test_gpu_parity.py runs in both); the single-pass 'tf32' figure (north_star's <= 1e-3 tolerance) is reported
under extra.tf32.

N > 1 (torchrun): C2 at batch 64 per GPU is "replicas only" (the 60 MB all-reduce of layer 0
costs more than the 0.1 ms step), so `value` is N independent replicas, one per GPU, no collective; the
batch-sharded data-parallel path (peer-memory reduce-scatter + update + all-gather) is measured beside it under
extra.dp: on C2 itself, and on the large-batch config where it pays (C5 layer 10000 -> 4096, CD-10, global batch
8192, strong scaling), after a parity check of the sharded update against the CPU oracle.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402

LAYERS = [10000, 1500, 500]
BATCH = 64
CD_K = 1
PARAMS = dict(LEARNING_RATE=0.1, WEIGHT_PENALTY=1e-4, INIT_MOMENTUM=0.5, FINAL_MOMENTUM=0.95,
              LEARNING_RATE_DYNAMIC=True, CD=CD_K, SPARSITY=False, SPARSITY_FACTOR=0.1)
N_DISTINCT_BATCHES = 64      # 64 x 2.56 MB of inputs rotate through the steps
METRIC = "cd_k_train_samples_per_s"
UNIT = "samples/s"
WORKLOAD = "C2 iDBN [10000,1500,500] CD-1 batch 64 per GPU (idbn.py:199-204)"
DTYPES = {"fp32": "f32", "tf32": "tf32",
          "tf32x2": "tf32x2 (tcgen05 kind::tf32 on hi + lo operand terms, fp32 accumulate: fp32-faithful)"}


def step_bytes(layers, k):
    """Algorithmic HBM bytes of one iDBN.train minibatch: per layer
    (4 + (1+2k)) * 4 * V * H for train_epoch + 4 * V * H for the post-update forward."""
    return sum((4 + (1 + 2 * k) + 1) * 4 * v * h for v, h in zip(layers[:-1], layers[1:]))


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def hbm_peak():
    d = load_peaks()
    if "hbm_gbs" in d:
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tf32_peak():
    d = load_peaks()
    if "bf16_tflops_sustained" in d:
        return float(d["bf16_tflops_sustained"]) / 2, "half of MEASURED_PEAKS.json bf16_tflops_sustained"
    return 1125.0, "fallback (nominal dense bf16 2250 / 2)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            text, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in text.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        return out


def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must use the box's cores regardless."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:                                   # noqa: BLE001
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU path for this workload: the oracle port (oracle/rbm_oracle.py restates
    imdbn/models/{rbm,idbn}.py op for op in torch-CPU fp32), all host threads.  One process (rank 0) regardless of N:
    the reference has no multi-device path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _all_host_threads()
    from oracle import rbm_oracle as O
    from oracle.philox import RandomField
    steps, warm = args.steps, args.warmup
    layers = [O.new_state(v, h, seed=i, lr=0.1, weight_decay=1e-4, momentum=0.5, final_momentum=0.95,
                          dynamic_lr=True) for i, (v, h) in enumerate(zip(LAYERS[:-1], LAYERS[1:]))]
    nb = min(N_DISTINCT_BATCHES, steps + warm)
    data = O.synthetic_images(nb * BATCH, LAYERS[0], seed=1234)
    t0 = None
    for i in range(warm + steps):
        if i == warm:
            t0 = time.perf_counter()
        x = data[(i % nb) * BATCH:(i % nb + 1) * BATCH]
        O.idbn_train_batch(layers, x, 0, CD_K, [RandomField(1, 2 * i), RandomField(2, 2 * i)])
    dt = time.perf_counter() - t0
    val = steps * BATCH / dt
    sample = f"{steps} minibatches of {BATCH} after {warm} warm-up, torch-CPU fp32, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH,
                       "note": "single process: the reference has no multi-device path"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


TRAFFIC_CSV = os.path.join("profiles", "r02_ncu_full_tc_stats.csv")


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel's layer-0 launch, from the committed
    `ncu --set full` summary (a cross-reference, not measured in this run); None if the file is absent."""
    import csv
    for rel in (TRAFFIC_CSV, os.path.join("profiles", "r01_ncu_full_tc_stats.csv")):
        path = os.path.join(REPO, rel)
        try:
            rows = list(csv.reader(open(path)))
            h, units = rows[0], rows[1]
            ir, iw, ig = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("launch__grid_size")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            best = max((r for r in rows[2:] if r and "k_tc_stats" in r[1]), key=lambda r: float(r[ig]))
            return float(best[ir]) * scale[units[ir]] + float(best[iw]) * scale[units[iw]], rel
        except Exception:                                   # noqa: BLE001
            continue
    return None, None


def cpu_baseline(seconds_budget=12.0):
    cores = _all_host_threads()
    from oracle import rbm_oracle as O
    from oracle.philox import RandomField
    layers = [O.new_state(v, h, seed=i, lr=0.1, weight_decay=1e-4, momentum=0.5, final_momentum=0.95,
                          dynamic_lr=True) for i, (v, h) in enumerate(zip(LAYERS[:-1], LAYERS[1:]))]
    data = O.synthetic_images(8 * BATCH, LAYERS[0], seed=1234)
    n, t0 = 0, None
    while True:
        if n == 2:
            t0 = time.perf_counter()
        x = data[(n % 8) * BATCH:(n % 8 + 1) * BATCH]
        O.idbn_train_batch(layers, x, 0, CD_K, [RandomField(1, 2 * n), RandomField(2, 2 * n)])
        n += 1
        if t0 is not None and (time.perf_counter() - t0 > seconds_budget or n >= 66):
            break
    dt = time.perf_counter() - t0
    steps = n - 2
    return {"value": steps * BATCH / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} minibatches of {BATCH} (C2 workload) after 2 warm-up, oracle port, "
                      f"torch-CPU fp32, {cores} threads of {os.cpu_count()} cpus"}


# ------------------------------------------------------------------------------------------------
def make_c2(M, dev, precision, warm, reserve, resident, seed=0):
    """Model + batch accessor of the device-resident C2 loop, warmed up."""
    M.set_precision(precision)
    torch.manual_seed(seed)
    model = M.iDBN(LAYERS, dict(PARAMS), None, None, dev)
    if reserve >= 0:
        model.pipeline_layers = True
        model.pipeline_reserve_sms = reserve

    def batch(i):
        return resident[i % N_DISTINCT_BATCHES]

    # untimed pre-warm: one pass over the distinct input buffers (workspace growth, tensor-map cache), then the W
    # declared warm-up steps
    for i in range(N_DISTINCT_BATCHES):
        model.train_step(batch(i), 0, 1, next_v=batch(i + 1))
    for i in range(warm):
        model.train_step(batch(i), 0, 1, next_v=batch(i + 1))
    model.sync()
    torch.cuda.synchronize()
    return model, batch


def run_ours(args):
    import torch.distributed as td
    import multimodal_idbn_b200 as M
    from multimodal_idbn_b200 import _lib as L

    rank = M.dist.init_from_env("nccl")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", torch.cuda.current_device())
    M.load_library()
    M.set_precision(args.precision)
    os.chdir(os.environ.get("TMPDIR", "/tmp"))       # iDBN creates logs-idbn/ in CWD (idbn.py:115)

    g = torch.Generator().manual_seed(1234 + rank)   # (replicas: every rank trains on its own stream of images)
    host = (torch.rand(N_DISTINCT_BATCHES, BATCH, LAYERS[0], generator=g) < 0.10).float().pin_memory()
    resident = host.to(dev)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    steps, warm = args.steps, args.warmup
    # clocks and throttle reasons: nvidia-smi polls from BEFORE the warm-up (its start-up enumerates every GPU of the
    # box and holds driver locks for a while: started at the edge of a 3 ms timed region it perturbs the launches, most
    # visibly with eight processes) until after the end-to-end loops
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    model, batch = make_c2(M, dev, args.precision, warm, args.pipeline_reserve, resident, seed=rank)
    # ---------------- device-resident timing: `value`
    barrier()
    l0 = M.total_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        # the loop knows its next minibatch (as iDBN.train does through the prefetcher)
        model.train_step(batch(warm + i), 0, 1, next_v=batch(warm + i + 1))
    model.sync()                                                   # (pipelined layers: join the side stream)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = M.total_launches() - l0
    value = world * BATCH * steps / (ms * 1e-3)

    # ---------------- end to end through the public API with host buffers: `e2e`
    loss_host = torch.full((steps + warm, len(model.layers)), float("nan")).pin_memory()

    def e2e_loop(n, off):
        # Host batches -> side-stream H2D with one batch of lookahead (prefetch_to_device, the loop iDBN.train
        # runs); the per-layer losses of every step reach the pinned host array `loss_host` through one 8-byte
        # device-to-host copy per step on the upper layers' stream (off layer 0's critical path).
        loader = [(host[(off + i) % N_DISTINCT_BATCHES],) for i in range(n)]
        cur, i = None, 0
        for b in M.prefetch_to_device(loader, dev):
            nxt = b[0]
            if cur is not None:
                model.train_step(cur, 0, 1, next_v=nxt, loss_out=loss_host[off + i])
                i += 1
            cur = nxt
        if cur is not None:
            model.train_step(cur, 0, 1, loss_out=loss_host[off + i])
        t_enq = time.perf_counter()
        model.sync()
        return t_enq

    # K steps, five times; the MEDIAN repetition is reported (this loop includes the host: a single repetition
    # is exposed to scheduling noise of the box), all five are listed in e2e.runs_ms
    e2e_loop(warm, 0)
    e2e_runs, e2e_host = [], []
    for _ in range(5):
        barrier()
        t0 = time.perf_counter()
        t_enq = e2e_loop(steps, warm)
        barrier()
        e2e_runs.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
        e2e_host.append((t_enq - t0) * 1e3)
    e2e_ms = sorted(e2e_runs)[2]
    clocks = sampler.stop() if sampler else None
    e2e_val = world * BATCH * steps / (e2e_ms * 1e-3)
    assert torch.isfinite(loss_host[warm:]).all()

    # ---------------- roofline of the dominant kernel (layer-0 statistics + update), CUDA events on ITS stream
    ctx = (model.__dict__.get("_fused") or {}).get("ctx0") or L.context_for(model.layers[0].W)[0]   # layer 0's context
    ctx.profile(True)
    n_prof = min(steps, 20)
    for i in range(n_prof):
        model.train_step(batch(i), 0, 1, next_v=batch(i + 1))
    model.sync()
    torch.cuda.synchronize()
    V0, H0 = LAYERS[0], LAYERS[1]
    kern = {}
    for name, kind in (("up", L.KERNEL_UP), ("down", L.KERNEL_DOWN), ("stats_update", L.KERNEL_STATS),
                       ("pack_operands", L.KERNEL_PACK)):
        tot, cnt = ctx.profile_read(kind, V0, H0)
        kern[name] = {"ms_avg": tot / max(1, cnt), "launches": cnt}
    ctx.profile(False)
    peak, peak_src = hbm_peak()
    upd_bytes = 4 * 4 * V0 * H0                      # read W, W_m; write W, W_m
    pass_bytes = 4 * V0 * H0
    kbytes = {"up": pass_bytes, "down": pass_bytes, "stats_update": upd_bytes, "pack_operands": None}
    ach = upd_bytes / (kern["stats_update"]["ms_avg"] * 1e-3) / 1e9 if kern["stats_update"]["ms_avg"] else None
    sb = step_bytes(LAYERS, CD_K)
    sb_moved = sb - 4 * V0 * H0          # the post-update forward of layer 0 is fused with the next positive phase
    step_gbs = sb / (ms / steps * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic_bytes()
    del model
    torch.cuda.empty_cache()

    extra = {}
    if args.extras:
        # ---------------- the same loop in single-pass tf32 (north_star's <= 1e-3 mode)
        if args.precision != "tf32":
            m2, b2 = make_c2(M, dev, "tf32", warm, args.pipeline_reserve, resident, seed=rank)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(steps):
                m2.train_step(b2(warm + i), 0, 1, next_v=b2(warm + i + 1))
            m2.sync()
            a1.record()
            barrier()
            ms2 = max_over_ranks(a0.elapsed_time(a1))
            extra["tf32"] = {"value": world * BATCH * steps / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / steps,
                             "note": "same workload, single-pass tcgen05 kind::tf32 (activations / weights within 1e-3 "
                                     "of fp32; sampled states are NOT bit-exact)"}
            del m2
            torch.cuda.empty_cache()
            M.set_precision(args.precision)
        # ---------------- second metric of BASELINE.json: cross-modal chain-steps/s (config C4)
        extra["c4_cross_modal"] = c4_metric(M, dev, 65536 // world, world)
        if world == 1:
            extra["c3_train_joint"] = c3_metric(M, dev, args.precision)
            extra["c5_large_batch"] = c5_metric(M, dev)
    if world > 1 and args.extras:
        extra["dp"] = dp_block(M, dev, rank, world, args.precision, barrier, max_over_ranks)

    if rank != 0:
        return
    cpu = cpu_baseline() if world == 1 and not args.no_cpu_baseline else None
    par = "single GPU" if world == 1 else (
        f"{world} independent replicas, one per GPU, no collective (SURVEY 8e row 3: C2 at batch 64 per GPU is "
        f"'replicas only' -- the exchange of layer 0's 60 MB statistic costs more than the step); the batch-sharded "
        f"data-parallel path is measured under extra.dp")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPES[args.precision], "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "global_batch": BATCH * world,
                   "parallelism": par,
                   "layer_pipelining": (f"upper layers of minibatch t next to layer 0 of minibatch t+1 on disjoint SM "
                                        f"partitions (CUDA green contexts, {args.pipeline_reserve} SMs for the upper "
                                        f"layers; iDBN.pipeline_layers; same arithmetic, split-K order follows the "
                                        f"partition sizes)"
                                        if args.pipeline_reserve >= 0 else "off"),
                   "precision_mode": args.precision,
                   "untimed_prewarm_steps": N_DISTINCT_BATCHES,
                   "prewarm_note": f"{N_DISTINCT_BATCHES} untimed steps (one per distinct input buffer: workspace growth, "
                                   f"tensor-map cache) precede the {warm} declared warm-up steps",
                   "l2": "state (W, W_m of both layers: 252 MB) + 164 MB of rotating inputs exceed the "
                         "126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms / steps, "runs_ms": e2e_runs,
                "host_enqueue_ms_per_step": sorted(e2e_host)[2] / steps,      # host time to enqueue the loop (rank 0)
                "h2d_bytes_per_step": BATCH * LAYERS[0] * 4, "d2h_bytes_per_step": 4 * len(LAYERS[1:]),
                "d2h": "per-layer losses of every step: one 8-byte device-to-host copy per step on the upper layers' stream into pinned host memory"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_tc_stats: layer-0 CD statistics + momentum/weight-decay update",
                     "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": (ach / peak) if ach else None, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": upd_bytes, "peak_source": peak_src,
                     "avg_launch_ms": kern["stats_update"]["ms_avg"]},
        "kernels": {k: dict(v, gbs=(kbytes[k] / (v["ms_avg"] * 1e-3) / 1e9 if kbytes[k] and v["ms_avg"] else None))
                    for k, v in kern.items()},
        "roofline_step": {"algorithmic_bytes": sb, "achieved_gbs": step_gbs, "frac": step_gbs / peak,
                          "bytes_moved": sb_moved, "moved_gbs": sb_moved / (ms / steps * 1e-3) / 1e9,
                          "moved_frac": sb_moved / (ms / steps * 1e-3) / 1e9 / peak,
                          "note": "algorithmic = SURVEY 8d (504 MB); moved = what the implementation streams (layer 0's "
                                  "post-update forward shares its pass over W with the next positive phase)"},
        "cpu_baseline": cpu,
        "extra": extra,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def _timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def c4_metric(M, dev, n_chains, world, steps=50, K=64):
    """Config C4 (imdbn.py:386-488) on this rank's share of the 65 536 chains, joint RBM (500+32) -> 256:
    IMG->TXT 50-step conditional Gibbs on `n_chains` items; TXT->IMG noisy mean-field annealing on `n_chains`
    independent chains; and iMDBN._cross_reconstruct with best-of-K = 64 (free-energy hook installed) on
    n_chains / 64 items (rows = items x candidates = n_chains).  Roofline: tensor pipe with the MINIMAL clamp-aware
    FLOPs of SURVEY 8d (4*Dz*H per TXT->IMG chain-step, 4*K*H per IMG->TXT chain-step), TF32 dense peak."""
    Dz, Kl, H = 500, 32, 256
    V = Dz + Kl
    torch.manual_seed(3)
    r = M.RBM(V, H, 0.04, 1e-4, 0.5, softmax_groups=[(Dz, V)]).to(dev)
    z = torch.rand(n_chains, Dz, device=dev)
    y = torch.nn.functional.one_hot(torch.randint(0, Kl, (n_chains,), device=dev), Kl).float()
    vk1 = torch.zeros(n_chains, V, device=dev); km1 = torch.zeros_like(vk1); vk1[:, :Dz] = z; km1[:, :Dz] = 1
    vk2 = torch.zeros(n_chains, V, device=dev); km2 = torch.zeros_like(vk2); vk2[:, Dz:] = y; km2[:, Dz:] = 1
    mu = torch.rand(n_chains, Dz, device=dev)
    peak, _ = tf32_peak()
    out = {"chains_this_rank": n_chains, "steps": steps, "precision": M.get_precision()}
    r._mu_pull = None
    ms = _timed(lambda: r.conditional_gibbs(vk1, km1, n_steps=steps, clamp_prefix=Dz), 3)
    cps = n_chains * steps / (ms * 1e-3)
    out["img2txt_cond_gibbs"] = {"chain_steps_per_s": cps * world, "ms": ms,
                                 "min_tflops": cps * 4 * Kl * H / 1e12, "tensor_frac": cps * 4 * Kl * H / 1e12 / peak,
                                 "engine": "k_label_gibbs: fp32 FFMA on the 32 x 256 label block (the clamped z enters "
                                           "once through one up-pass GEMM); tensor_frac is against the TF32 peak for scale"}
    r._mu_pull = {"mu_k": mu, "eta0": 0.15}
    prev = M.get_precision()
    out["txt2img_noisy_mf"] = {}
    for mode in dict.fromkeys(("tf32", prev)):
        M.set_precision(mode)
        ms = _timed(lambda: r.noisy_meanfield_annealed(vk2, km2, n_steps=steps, clamp_suffix=Dz), 3)
        cps = n_chains * steps / (ms * 1e-3)
        out["txt2img_noisy_mf"][mode] = {
            "chain_steps_per_s": cps * world, "ms": ms, "min_tflops": cps * 4 * Dz * H / 1e12,
            "tensor_frac": cps * 4 * Dz * H / 1e12 / peak,
            "engine": ("k_chain_tc: ONE persistent tcgen05 kernel, 48 chains per CTA, state in shared memory as the MMA operand"
                       if mode == "tf32" else "stepped: two k_tc_stream GEMMs + finish kernels per step (exact mode)")}
    M.set_precision(prev)
    out["txt2img_noisy_mf"]["bound"] = ("instruction issue, not the tensor pipe: per chain-step (H + Dz) = 756 Gaussian draws "
                                       "(Philox4x32-10 + Box-Muller) and sigmoids, ~1e5 thread instructions")
    r._mu_pull = None
    # best-of-K through the product API
    items = max(1, n_chains // K)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = M.iMDBN([1000, 500], H, params=dict(PARAMS, JOINT_LEARNING_RATE=0.04, CROSS_GIBBS_STEPS=steps),
                    dataloader=None, val_loader=None, device=dev, num_labels=Kl)
    m.N_CANDIDATES = K
    m.z_class_mean = torch.rand(Kl, Dz, device=dev)
    type(m.joint_rbm).free_energy = M.rbm_free_energy
    try:
        zi, yi = z[:items].contiguous(), y[:items].contiguous()
        ms = _timed(lambda: m._cross_reconstruct(zi, yi, steps=steps), 2)
    finally:
        del type(m.joint_rbm).free_energy
    sweeps = items * (steps + 1 + steps + (K - 1))
    out["cross_reconstruct_best_of_k"] = {"items": items, "K": K, "ms": ms, "items_per_s": items / (ms * 1e-3) * world,
                                          "sweeps_per_s": sweeps / (ms * 1e-3) * world,
                                          "includes": "IMG->TXT 50+1 sweeps, TXT->IMG 50 sweeps + 63 refinements, 64 free "
                                                      "energies, argmin, decode 500 -> 1000"}
    return out


def c3_metric(M, dev, precision, n_batches=4):
    """Config C3: one main-phase batch of iMDBN.train_joint (imdbn.py:553-639): represent + free CD-1 on the joint RBM
    + aux clamped CD (30 cond steps) + _cross_reconstruct (50 steps) + metrics, batch 64; CPU port timed beside it."""
    from oracle import rbm_oracle as O
    from oracle.philox import RandomField
    P = dict(PARAMS, JOINT_LEARNING_RATE=0.04, JOINT_CD=1, CROSS_GIBBS_STEPS=50, JOINT_AUX_COND_STEPS=30)
    N, B = 64 * n_batches, 64
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(N, 10000, generator=g) < 0.1).float()
    y = torch.nn.functional.one_hot(torch.randint(0, 32, (N,), generator=g), 32).float()
    dl = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x.pin_memory(), y.pin_memory()), batch_size=B)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = M.iMDBN([10000, 1500, 500], 256, params=P, dataloader=dl, val_loader=None, device=dev, num_labels=32)
        m.WARMUP_Y_EPOCHS = 0            # time the main phase
        m.train_joint(1)                 # warm-up epoch
        torch.cuda.synchronize()
        l0 = M.total_launches()
        t0 = time.perf_counter()
        m.train_joint(3)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        launches = M.total_launches() - l0
    nb = 3 * n_batches
    out = {"ms_per_batch": dt / nb * 1e3, "samples_per_s": nb * B / dt, "kernel_launches_per_batch": launches / nb,
           "batch": B, "precision": precision, "timing": "wall clock over 3 epochs x 4 batches, includes the host"}
    # CPU port of the same batch (one batch, bounded)
    _all_host_threads()
    img_layers = [O.new_state(v, h, seed=i) for i, (v, h) in enumerate([(10000, 1500), (1500, 500)])]
    joint = O.new_state(532, 256, seed=7, groups=[(500, 532)], lr=0.04, weight_decay=1e-4, momentum=0.5,
                        final_momentum=0.95, dynamic_lr=True)
    xb, yb = x[:B], y[:B]
    zc = torch.rand(32, 500)
    t0 = time.perf_counter()
    z = O.idbn_represent(img_layers, xb)
    vplus = torch.cat([z, yb], 1)
    O.cd_train(joint, vplus, 9, 1, RandomField(1, 0))
    vk = torch.zeros(B, 532); km = torch.zeros(B, 532); vk[:, 500:] = yb; km[:, 500:] = 1
    O.cd_train_clamped(joint, vk, km, 9, k=1, cond_init_steps=30, sample_h=False, sample_v=False,
                       reclamp_negative=False, aux_lr_mult=0.3, use_noisy_init=True, fld=RandomField(1, 1))
    O.cross_reconstruct(img_layers, joint, z, yb, 50, z_class_mean=zc, fld_i2t=RandomField(1, 2),
                        fld_t2i=RandomField(1, 3), fld_refine=[RandomField(1, 4 + c) for c in range(4)])
    out["cpu_port_ms_per_batch"] = (time.perf_counter() - t0) * 1e3
    out["cpu_cores"] = torch.get_num_threads()
    return out


def c5_metric(M, dev):
    """Config C5: CD-10 train_epoch per layer of the widened stack [10000,4096,2048] + joint (2048+32) -> 1024 at batch
    64 ... 8192, single-pass tf32 (the tensor-bound mode); FLOPs = (3 + 2k) * 2 * B * V * H against the TF32
    dense peak."""
    prev = M.get_precision()
    M.set_precision("tf32")
    peak, peak_src = tf32_peak()
    k = 10
    out = {"precision": "tf32", "cd_k": k, "tf32_peak_tflops": peak, "peak_source": peak_src, "layers": {}}
    try:
        for name, V, H, groups in (("L0 10000->4096", 10000, 4096, None), ("L1 4096->2048", 4096, 2048, None),
                                   ("joint 2080->1024", 2080, 1024, [(2048, 2080)])):
            r = M.RBM(V, H, 0.1, 1e-4, 0.5, softmax_groups=groups).to(dev)
            rows = {}
            for B in (64, 512, 4096, 8192):
                if groups:
                    x = torch.cat([(torch.rand(B, 2048, device=dev) < 0.3).float(),
                                   torch.nn.functional.one_hot(torch.randint(0, 32, (B,), device=dev), 32).float()], 1)
                else:
                    x = (torch.rand(B, V, device=dev) < 0.1).float()
                ms = _timed(lambda: r.train_epoch(x, 0, 1, CD=k), 2 if B >= 4096 else 4)
                tf = (3 + 2 * k) * 2.0 * B * V * H / (ms * 1e-3) / 1e12
                rows[str(B)] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "tflops": tf, "frac": tf / peak}
                del x
            out["layers"][name] = rows
            del r
            torch.cuda.empty_cache()
        # whole three-RBM step at batch 8192: sum of the three layers
        tot = sum(out["layers"][n]["8192"]["ms_per_step"] for n in out["layers"])
        fl = 23 * 2.0 * 8192 * (10000 * 4096 + 4096 * 2048 + 2080 * 1024)
        out["stack_batch_8192"] = {"ms_per_step": tot, "samples_per_s": 8192 / (tot * 1e-3),
                                   "tflops": fl / (tot * 1e-3) / 1e12, "frac": fl / (tot * 1e-3) / 1e12 / peak}
    finally:
        M.set_precision(prev)
    return out


def dp_block(M, dev, rank, world, precision, barrier, max_over_ranks):
    """Batch-sharded data parallelism (dist.py: statistics exchanged over NVLink peer memory, reduce-scatter + slab update
    + all-gather in one kernel): (1) parity of one sharded CD-1 update of the C2 bottom layer against the CPU oracle of
    the whole global batch, (2) C2 weak scaling with the exchange, (3) the large-batch config where sharding pays:
    RBM 10000 -> 4096, CD-10, GLOBAL batch 8192 (strong scaling), single-pass tf32."""
    import torch.distributed as td
    from oracle import rbm_oracle as O
    from oracle.philox import RandomField
    out = {}
    M.dist.enable()
    ds = M.dist.state()
    out["exchange"] = ("NVLink peer memory" + (" with NVSwitch multimem reduce / broadcast" if ds.multicast else "")) \
        if ds.p2p else "NCCL all-reduce"
    # ---- (1) parity vs the oracle
    M.set_precision(precision)
    V, H, Bg = LAYERS[0], LAYERS[1], BATCH * world
    st = O.new_state(V, H, seed=5, lr=0.1, weight_decay=1e-4, momentum=0.5, final_momentum=0.95, dynamic_lr=True)
    r = M.RBM(V, H, 0.1, 1e-4, 0.5, dynamic_lr=True, final_momentum=0.95).to(dev)
    with torch.no_grad():
        r.W.data.copy_(st.W)
    data = O.synthetic_images(Bg, V, seed=77)
    lo, hi = M.dist.shard_rows(Bg, rank, world)
    r.set_rng(41, 0)
    loss = r.train_epoch(data[lo:hi].to(dev), 0, 1, CD=1)
    torch.cuda.synchronize()
    if rank == 0:
        _all_host_threads()
        W0 = st.W.clone()
        loss_ref, _ = O.cd_train(st, data, 0, 1, RandomField(41, 0))
        dW_ref, dW = st.W - W0, r.W.detach().cpu() - W0
        err = float((dW - dW_ref).abs().max())
        bad = float(((dW - dW_ref).abs() > 2e-7 + 1e-3 * dW_ref.abs()).float().mean())
        out["parity_vs_oracle"] = {"shape": f"{V}->{H}", "global_batch": Bg, "precision": precision,
                                   "max_abs_dW_err": err, "max_abs_dW": float(dW_ref.abs().max()),
                                   "frac_elements_off_1e-3": bad, "loss": float(loss), "loss_ref": float(loss_ref)}
        tol = 2e-2 if precision == "tf32" else 5e-3
        assert abs(float(loss) - float(loss_ref)) <= 1e-3 * abs(float(loss_ref)) + 1e-6, "DP loss differs from the oracle"
        assert bad < tol, f"DP update differs from the oracle ({bad})"
    w = r.W.detach().clone()
    td.broadcast(w, 0)
    assert torch.equal(w, r.W.detach()), "replicas diverged"
    del r
    # ---- (2) C2 with the exchange, weak scaling
    torch.manual_seed(0)
    model = M.iDBN(LAYERS, dict(PARAMS), None, None, dev)
    g = torch.Generator().manual_seed(99 + rank)
    xs = (torch.rand(8, BATCH, V, generator=g) < 0.1).float().to(dev)
    for i in range(6):
        model.train_step(xs[i % 8], 0, 1, next_v=xs[(i + 1) % 8])
    barrier()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        model.train_step(xs[i % 8], 0, 1, next_v=xs[(i + 1) % 8])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / n
    out["c2_sharded"] = {"global_batch": Bg, "ms_per_step": ms, "samples_per_s": Bg / (ms * 1e-3),
                         "note": "every step exchanges the 60 MB layer-0 statistic: the config SURVEY 8e calls pointless to shard"}
    del model, xs
    torch.cuda.empty_cache()
    # ---- (3) large batch, strong scaling
    M.set_precision("tf32")
    Vl, Hl, Bgl, k = 10000, 4096, 8192, 10
    torch.manual_seed(1)
    rl = M.RBM(Vl, Hl, 0.1, 1e-4, 0.5).to(dev)
    xl = (torch.rand(Bgl // world, Vl, device=dev) < 0.1).float()
    for _ in range(2):
        rl.train_epoch(xl, 0, 1, CD=k)
    barrier()
    n = 3
    e0.record()
    for _ in range(n):
        rl.train_epoch(xl, 0, 1, CD=k)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / n
    fl = (3 + 2 * k) * 2.0 * Bgl * Vl * Hl
    out["large_batch_strong"] = {"workload": f"RBM {Vl}->{Hl} CD-{k}, global batch {Bgl} ({Bgl // world} rows per GPU), tf32",
                                 "ms_per_step": ms, "samples_per_s": Bgl / (ms * 1e-3),
                                 "tflops_total": fl / (ms * 1e-3) / 1e12,
                                 "note": "compare with extra.c5_large_batch.layers['L0 10000->4096']['8192'] of the 1-GPU run"}
    del xl
    torch.cuda.empty_cache()
    # ---- (4) the same layer, weak scaling: 8192 rows per GPU (global batch 8192 x N)
    xw = (torch.rand(Bgl, Vl, device=dev) < 0.1).float()
    for _ in range(2):
        rl.train_epoch(xw, 0, 1, CD=k)
    barrier()
    n = 2
    e0.record()
    for _ in range(n):
        rl.train_epoch(xw, 0, 1, CD=k)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / n
    out["large_batch_weak"] = {"workload": f"RBM {Vl}->{Hl} CD-{k}, {Bgl} rows per GPU (global batch {Bgl * world}), tf32",
                               "ms_per_step": ms, "samples_per_s": Bgl * world / (ms * 1e-3),
                               "tflops_total": fl * world / (ms * 1e-3) / 1e12}
    del rl, xw
    torch.cuda.empty_cache()
    M.dist.disable()
    M.set_precision(precision)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tf32x2", choices=["fp32", "tf32", "tf32x2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the extra.* blocks (C3 / C4 / C5 / dp)")
    ap.add_argument("--pipeline-reserve", type=int, default=16,
                    help="SMs left to the upper layers, which then run on a side stream concurrently with the next "
                         "layer-0 update (-1 = layers run back to back on one stream)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        if args.steps > 64:
            args.steps = 64            # bounded sample: ~0.06 s per CPU step
        args.warmup = min(args.warmup, 2)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
