#!/usr/bin/env python
"""bench.py -- CD-k training throughput of the iDBN hot loop on B200 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|tf32]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" = one minibatch of ``iDBN.train`` on [10000,1500,500], CD-1, batch 64 per GPU
(reference idbn.py:199-204): for each layer one ``train_epoch`` and one ``forward`` with the updated
weights.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
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


def step_bytes(layers, k):
    """Algorithmic HBM bytes of one iDBN.train minibatch: per layer
    (4 + (1+2k)) * 4 * V * H for train_epoch + 4 * V * H for the post-update forward."""
    return sum((4 + (1 + 2 * k) + 1) * 4 * v * h for v, h in zip(layers[:-1], layers[1:]))


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU path for this workload: the oracle port (oracle/rbm_oracle.py restates
    imdbn/models/{rbm,idbn}.py op for op in torch-CPU fp32), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
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
    cores = torch.get_num_threads()
    sample = f"{steps} minibatches of {BATCH} after {warm} warm-up, torch-CPU fp32, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2 iDBN [10000,1500,500] CD-1 batch 64 (idbn.py:199-204)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel's layer-0 launch, from the committed
    `ncu --set full` summary (profiles/r01_ncu_full_tc_stats.csv); None if the file is absent."""
    import csv
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_full_tc_stats.csv")
    try:
        rows = list(csv.reader(open(path)))
        h, units = rows[0], rows[1]
        ir, iw, ig = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("launch__grid_size")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        best = max((r for r in rows[2:] if r and "k_tc_stats" in r[1]), key=lambda r: float(r[ig]))
        return float(best[ir]) * scale[units[ir]] + float(best[iw]) * scale[units[iw]]
    except Exception:                                   # noqa: BLE001
        return None


def cpu_baseline(seconds_budget=15.0):
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
    cores = torch.get_num_threads()
    return {"value": steps * BATCH / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} minibatches of {BATCH} (C2 workload) after 2 warm-up, oracle port, "
                      f"torch-CPU fp32, {cores} threads of {os.cpu_count()} cpus"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as td
    import multimodal_idbn_b200 as M
    from multimodal_idbn_b200 import _lib as L

    rank = M.dist.init_from_env("nccl")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", torch.cuda.current_device())
    M.load_library()
    M.set_precision(args.precision)
    if world > 1:
        M.dist.enable()

    torch.manual_seed(0)
    os.chdir(os.environ.get("TMPDIR", "/tmp"))       # iDBN creates logs-idbn/ in CWD (idbn.py:115)
    model = M.iDBN(LAYERS, dict(PARAMS), None, None, dev)
    if args.pipeline_reserve >= 0 and world == 1:
        model.pipeline_layers = True
        model.pipeline_reserve_sms = args.pipeline_reserve
    g = torch.Generator().manual_seed(1234)
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

    dp_desc = "single GPU"
    if world > 1:
        ds = M.dist.state()
        if ds.p2p:
            mc = ds.multicast
            dp_desc = (f"dp{world} (batch-sharded; statistics exchanged over NVLink peer memory: reduce-scatter + slab update "
                       f"+ all-gather in one kernel" + (", NVSwitch multimem reduce / broadcast when available)" if mc else ")"))
        else:
            dp_desc = f"dp{world} (batch-sharded, NCCL all-reduce of dS)"
    steps, warm = args.steps, args.warmup
    # ---------------- device-resident timing: `value`
    def batch(i):
        return resident[i % N_DISTINCT_BATCHES]

    # untimed setup: one pass over the distinct input buffers fills the library's TMA-descriptor cache (descriptors
    # are keyed by buffer address; encoding them costs host time the first time a buffer is seen), then the W
    # warm-up steps proper
    for i in range(N_DISTINCT_BATCHES):
        model.train_step(batch(i), 0, 1, next_v=batch(i + 1))
    for i in range(warm):
        model.train_step(batch(i), 0, 1, next_v=batch(i + 1))
    barrier()
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
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
    clocks = sampler.stop() if sampler else None
    value = world * BATCH * steps / (ms * 1e-3)

    # ---------------- end to end through the public API with host buffers: `e2e`
    loss_host = torch.full((steps + warm, len(model.layers)), float("nan")).pin_memory()

    def e2e_loop(n, off):
        # Host batches -> side-stream H2D with one batch of lookahead (prefetch_to_device, the loop iDBN.train
        # runs); the per-layer losses of every step are written by the update kernels straight into the pinned
        # host array `loss_host` (mapped memory: the device-to-host read-back is a PCIe store, no copy node).
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

    # K steps, three times; the MEDIAN repetition is reported (this loop includes the host: a single repetition
    # is exposed to scheduling noise of the box), all three are listed in e2e.runs_ms
    e2e_loop(warm, 0)
    e2e_runs = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        e2e_loop(steps, warm)
        barrier()
        e2e_runs.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
    e2e_ms = sorted(e2e_runs)[1]
    e2e_val = world * BATCH * steps / (e2e_ms * 1e-3)
    assert torch.isfinite(loss_host[warm:]).all()

    # ---------------- roofline of the dominant kernel (layer-0 statistics + update), CUDA events
    ctx = (model.__dict__.get("_fused") or {}).get("ctx0") or L.context_for(model.layers[0].W)[0]   # layer 0's context
    ctx.profile(True)
    n_prof = min(steps, 20)
    for i in range(n_prof):
        model.train_step(batch(i), 0, 1, next_v=batch(i + 1))
    torch.cuda.synchronize()
    V0, H0 = LAYERS[0], LAYERS[1]
    kern = {}
    for name, kind in (("up", L.KERNEL_UP), ("down", L.KERNEL_DOWN), ("stats_update", L.KERNEL_STATS)):
        tot, cnt = ctx.profile_read(kind, V0, H0)
        kern[name] = {"ms_avg": tot / max(1, cnt), "launches": cnt}
    ctx.profile(False)
    peak, peak_src = load_peaks()
    upd_bytes = 4 * 4 * V0 * H0                      # read W, W_m; write W, W_m
    pass_bytes = 4 * V0 * H0
    ach = upd_bytes / (kern["stats_update"]["ms_avg"] * 1e-3) / 1e9 if kern["stats_update"]["ms_avg"] else None
    sb = step_bytes(LAYERS, CD_K)
    step_gbs = sb / (ms / steps * 1e-3) / 1e9

    # ---------------- second metric of BASELINE.json: cross-modal chain-steps/s (config C4 shapes,
    # this rank's share of the 65 536 chains: joint RBM (500+32)->256, 50 steps, both directions)
    chains = chain_steps_metric(M, dev, 65536 // max(1, world) if world > 1 else 8192)
    large = large_batch_metric(M, dev) if rank == 0 and world == 1 else None

    if rank != 0:
        return
    cpu = cpu_baseline() if world == 1 and not args.no_cpu_baseline else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x2": "tf32x2 (split tf32 operands, fp32 accumulate: fp32-faithful)"}[args.precision], "data": "synthetic",
        "config": {"workload": "C2 iDBN [10000,1500,500] CD-1 batch 64 per GPU (idbn.py:199-204)",
                   "global_batch": BATCH * world,
                   "parallelism": dp_desc,
                   "layer_pipelining": (f"upper layers of minibatch t next to layer 0 of minibatch t+1 on disjoint SM "
                                        f"partitions (CUDA green contexts, {args.pipeline_reserve} SMs for the upper "
                                        f"layers; iDBN.pipeline_layers; same arithmetic, split-K order follows the "
                                        f"partition sizes)"
                                        if args.pipeline_reserve >= 0 and world == 1 else "off"),
                   "precision_mode": args.precision,
                   "l2": "state (W, W_m of both layers: 252 MB) + 164 MB of rotating inputs exceed the "
                         "126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms / steps, "runs_ms": e2e_runs,
                "h2d_bytes_per_step": BATCH * LAYERS[0] * 4, "d2h_bytes_per_step": 4 * len(model.layers),
                "d2h": "per-layer losses stored by the update kernels into mapped pinned host memory"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "layer-0 CD statistics + momentum/weight-decay update",
                     "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": (ach / peak) if ach else None, "traffic": ncu_traffic_bytes(),
                     "algorithmic_bytes_per_launch": upd_bytes, "peak_source": peak_src,
                     "avg_launch_ms": kern["stats_update"]["ms_avg"]},
        "kernels": {k: dict(v, gbs=(pass_bytes if k != "stats_update" else upd_bytes) /
                            (v["ms_avg"] * 1e-3) / 1e9 if v["ms_avg"] else None) for k, v in kern.items()},
        "roofline_step": {"algorithmic_bytes": sb, "achieved_gbs": step_gbs, "frac": step_gbs / peak},
        "cpu_baseline": cpu,
        "extra": {"chain_steps_per_s": chains, "large_batch": large},
    }
    print(json.dumps(line), flush=True)


def chain_steps_metric(M, dev, n_chains, steps=50):
    """IMG->TXT conditional Gibbs and TXT->IMG noisy mean-field annealing (iMDBN._cross_reconstruct,
    imdbn.py:419-449) on `n_chains` chains; device-timed, chain-steps/s per direction (this rank)."""
    Dz, K, H = 500, 32, 256
    V = Dz + K
    torch.manual_seed(3)
    r = M.RBM(V, H, 0.04, 1e-4, 0.5, softmax_groups=[(Dz, V)]).to(dev)
    z = torch.rand(n_chains, Dz, device=dev)
    y = torch.nn.functional.one_hot(torch.randint(0, K, (n_chains,), device=dev), K).float()
    vk1 = torch.zeros(n_chains, V, device=dev); km1 = torch.zeros_like(vk1); vk1[:, :Dz] = z; km1[:, :Dz] = 1
    vk2 = torch.zeros(n_chains, V, device=dev); km2 = torch.zeros_like(vk2); vk2[:, Dz:] = y; km2[:, Dz:] = 1
    mu = torch.rand(n_chains, Dz, device=dev)
    out = {"chains": n_chains, "steps": steps}
    for name, fn in (("img2txt_cond_gibbs", lambda: r.conditional_gibbs(vk1, km1, n_steps=steps, clamp_prefix=Dz)),
                     ("txt2img_noisy_mf", lambda: r.noisy_meanfield_annealed(vk2, km2, n_steps=steps, clamp_suffix=Dz))):
        r._mu_pull = {"mu_k": mu, "eta0": 0.15} if name.startswith("txt2img") else None
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record(); torch.cuda.synchronize()
        out[name] = n_chains * steps * 3 / (e0.elapsed_time(e1) * 1e-3)
    r._mu_pull = None
    return out


def large_batch_metric(M, dev, B=8192, V=10000, H=4096, steps=4):
    """Tensor-bound end of the sweep (BASELINE config C5, widened layer 10000 -> 4096): CD-1 updates at batch
    `B` in tf32 mode; FLOPs = (3 + 2k) * 2 * B * V * H; the roofline is the TF32 dense peak = half the
    measured bf16 figure of MEASURED_PEAKS.json (sustained: the kernels run back to back for ~0.2 s)."""
    r = M.RBM(V, H, 0.1, 1e-4, 0.5).to(dev)
    x = (torch.rand(B, V, device=dev) < 0.1).float()
    for _ in range(2):
        r.train_epoch(x, 0, 1, CD=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r.train_epoch(x, 0, 1, CD=1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tflops = 5 * 2.0 * B * V * H / (ms * 1e-3) / 1e12
    peak = None
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["bf16_tflops_sustained"]) / 2
    except Exception:                                   # noqa: BLE001
        pass
    del r, x
    torch.cuda.empty_cache()
    return {"workload": f"RBM {V}->{H} CD-1 batch {B}, tf32", "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3),
            "tflops": tflops, "tf32_peak_tflops": peak, "frac": tflops / peak if peak else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # defaults: long enough for the steady state -- with the layers pipelined the step is short enough (~115 us)
    # that the host's first few hundred enqueues (cold caches, CPU clock ramp) would otherwise set the pace
    # (the warm-up also has to outlast the start-up regime in which the upper layer still lags behind the bottom one
    # and the bottom layer runs at its own, faster, pace: ~500 steps, tools/step_trend.py)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=600)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tf32x2", choices=["fp32", "tf32", "tf32x2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline-reserve", type=int, default=16,
                    help="SMs left to the upper layers, which then run on a side stream concurrently with the next "
                         "layer-0 update (-1 = layers run back to back on one stream)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        if args.steps > 64:
            args.steps = 64            # bounded sample: ~0.13 s per CPU step
        args.warmup = min(args.warmup, 2)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
