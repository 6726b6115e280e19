"""bench.py -- optimizer steps/s of the DotaClient optimizer hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c1]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, NCCL)

A "step" is one ``DotaOptimizer.train()`` call (forward, PPO loss, backward, gradient all-reduce, clip, Adam)
on one synthetic experience batch.  Default workload = BASELINE.json configs[1] ("c2": batch 256 x seq 512,
hidden 128, LSTM) PER GPU (weak scaling); ``value`` counts one such batch per GPU per step.

Keys: value (inputs resident in HBM), e2e (inputs in pinned host memory, H2D + result D2H inside the timed
region), roofline (recurrence fwd+bwd kernels: algorithmic bytes / CUDA-event time / measured HBM peak),
cpu_baseline (the oracle port of the reference's train() on the host cores, bounded sample), clocks,
gpu_launches.  ``--impl reference`` times the reference's CPU implementation (oracle port; the reference is
pure Python + torch CPU and cannot travel to the GPU box) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {                       # BASELINE.json configs; batch is PER GPU
    "c1": dict(batch=1, seq_len=64, hidden=128, cell="lstm"),
    "c2": dict(batch=256, seq_len=512, hidden=128, cell="lstm"),
    "c3": dict(batch=512, seq_len=512, hidden=256, cell="lstm"),
    "c4": dict(batch=512, seq_len=1024, hidden=512, cell="lstm"),
    # configs[4]: 40-agent replay stream through the in-process broker, the reference's own defaults
    # (optimizer.py:781-786: min_seq_per_epoch 1024, seq_len 16, epochs 4; policy.py:66: 256-wide GRU)
    "c5": dict(batch=1024, seq_len=16, hidden=256, cell="gru", stream=True),
}
FALLBACK_HBM_GBS = 6650.0         # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--config", choices=sorted(CONFIGS), default="c2")
    p.add_argument("--cell", choices=["gru", "lstm"], default=None)
    p.add_argument("--batch", type=int, default=None)
    p.add_argument("--seq-len", type=int, default=None)
    p.add_argument("--hidden", type=int, default=None)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cuda-profiler", action="store_true",
                   help="bracket the HBM-resident timed region with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    p.add_argument("--skip-e2e", action="store_true", help="(profiling only) skip the host-buffer timed region")
    return p.parse_args()


def resolve_config(args):
    cfg = dict(CONFIGS[args.config])
    for k, a in (("batch", args.batch), ("seq_len", args.seq_len), ("hidden", args.hidden), ("cell", args.cell)):
        if a is not None:
            cfg[k] = a
    return cfg


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU baseline
def log(msg):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.stderr.write("[bench %.1fs] %s\n" % (time.perf_counter() - _T0, msg))
    sys.stderr.flush()


_T0 = time.perf_counter()


def usable_cores():
    """Cores this process may actually use: affinity mask, capped by the cgroup CPU quota if there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


def cpu_reference_steps_per_sec(cfg, budget_s=20.0, threads=None):
    """The oracle port of the reference's ``DotaOptimizer.train`` (optimizer.py:581-689) on the host cores.

    Bounded sample: ``b`` sequences of the config's seq_len/hidden/cell; ``b`` starts at 1 and doubles until one
    step takes >= 1 s (or b reaches the config's batch), then steps are timed for ~budget_s.  steps/s for the full
    batch is extrapolated linearly in the batch size (the work is per-token).
    """
    import torch
    from oracle import ref_optimizer as RO
    from oracle.ref_policy import RefPolicy
    from dotaclient_b200.synthetic import make_rollout
    cores = threads or min(usable_cores(), 64)            # torch CPU stops scaling on these GEMM sizes long before 64
    torch.set_num_threads(cores)
    S, H, cell, B = cfg["seq_len"], cfg["hidden"], cfg["cell"], cfg["batch"]
    torch.manual_seed(7)
    opt = RO.RefOptimizer(RefPolicy(H, cell), seq_len=S)
    seqs = opt.experiences_from_rollout(make_rollout(S, 7))
    opt.train(seqs)                                        # untimed warm-up (a cold first call must not end the sizing loop)
    b = 1
    t_begin = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        opt.train(seqs)
        dt = time.perf_counter() - t0
        if dt >= 1.0 or b >= B or b >= 32 or time.perf_counter() - t_begin > budget_s / 2:
            break
        for i in range(b):
            seqs.extend(opt.experiences_from_rollout(make_rollout(S, 7 + b + i)))
        b *= 2
    t0, n = time.perf_counter(), 0
    while True:
        opt.train(seqs)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s / 2 or n >= 20:
            break
    b = len(seqs)
    sample_steps_per_s = n / el
    full = sample_steps_per_s * b / B
    sample = "oracle port of optimizer.py:581-689, batch %d x seq %d (hidden %d, %s), %d steps in %.1f s on %d threads; " \
             "scaled x%d/%d to batch %d" % (b, S, H, cell, n, el, cores, b, B, B)
    return full, cores, sample


def run_reference_arm(args, cfg):
    """``--impl reference``: the reference's CPU ``train()`` (oracle port: the reference is pure Python + torch CPU and
    cannot travel to the GPU box) on the host cores, same config/metric.  Each step is one train() on a bounded sample
    (``b`` of the config's ``batch`` sequences, full seq_len); steps/s is scaled by b/batch.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import ref_optimizer as RO
    from oracle.ref_policy import RefPolicy
    from dotaclient_b200.synthetic import make_rollout
    cores = min(usable_cores(), 64)
    torch.set_num_threads(cores)
    S, H, cell, B = cfg["seq_len"], cfg["hidden"], cfg["cell"], cfg["batch"]
    torch.manual_seed(7)
    opt = RO.RefOptimizer(RefPolicy(H, cell), seq_len=S)
    seqs = opt.experiences_from_rollout(make_rollout(S, 7))
    opt.train(seqs)                                        # untimed: thread pool / allocator warm-up must not end the sizing loop
    b = 1
    while True:                                            # grow the sample until one step takes >= 1 s (cap 32 sequences)
        t0 = time.perf_counter()
        opt.train(seqs)
        if time.perf_counter() - t0 >= 1.0 or b >= B or b >= 32:
            break
        for i in range(b):
            seqs.extend(opt.experiences_from_rollout(make_rollout(S, 7 + b + i)))
        b *= 2
    b = len(seqs)
    for _ in range(args.warmup):
        opt.train(seqs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        opt.train(seqs)
    el = time.perf_counter() - t0
    value = args.steps / el * b / B
    sample = "oracle port of optimizer.py:581-689, batch %d x seq %d (hidden %d, %s), %d steps in %.1f s on %d threads; " \
             "scaled x%d/%d to batch %d" % (b, S, H, cell, args.steps, el, cores, b, B, B)
    line = {
        "impl": "reference", "metric": "optimizer_steps_per_sec", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args.gpus, "cpu"),
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "one host: N reference ranks would share these cores, so the per-rank-batch rate does not grow with N",
    }
    emit(line)


def workload_config(cfg, n_gpus, where):
    return {"workload": "BASELINE configs: synthetic experience batch=%d seq=%d hidden=%d per GPU, %s cell; "
                        "one DotaOptimizer.train() per step" % (cfg["batch"], cfg["seq_len"], cfg["hidden"], cfg["cell"]),
            "batch_per_gpu": cfg["batch"], "global_batch": cfg["batch"] * n_gpus, "seq_len": cfg["seq_len"],
            "hidden": cfg["hidden"], "cell": cfg["cell"], "parallelism": "dp%d" % n_gpus,
            "l2": "per-step inputs exceed the 126 MB L2" if cfg["batch"] * cfg["seq_len"] * 2100 > 126e6
                  else "L2 flushed between steps", "where": where}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_stream(args, cfg):
    """BASELINE configs[4]: 40 agents (threads) publish pickled rollouts of 1000-1400 steps into the in-process
    ``MessageQueue``; every rank runs ``DotaOptimizer.run_iteration`` (pull -> prep -> epochs x train -> publish) and
    steady-state optimizer steps/s is reported.  Secondary configuration (not the headline bench line)."""
    import pickle
    import random
    import torch
    import torch.distributed as dist
    from dotaclient_b200.optimizer import DotaOptimizer, MessageQueue
    from dotaclient_b200.synthetic import make_rollout
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl")
    epochs = 4
    opt = DotaOptimizer(rmq_host="stream", rmq_port=rank, epochs=epochs, min_seq_per_epoch=cfg["batch"], seq_len=cfg["seq_len"],
                        learning_rate=5e-5, checkpoint=False, pretrained_model=None, mq_prefetch_count=1,
                        log_dir=tempfile.mkdtemp(), entropy_coef=5e-4, vf_coef=0.5, run_local=True,
                        hidden_size=cfg["hidden"], cell=cfg["cell"])
    agents = max(1, 40 // world)
    rng = random.Random(7 + rank)
    pool = [pickle.dumps(make_rollout(rng.randint(1000, 1400), 7 + 1000 * rank + i, with_canvas=True)) for i in range(8)]
    stop = threading.Event()
    mq = MessageQueue(host="stream", port=rank, prefetch_count=1, use_model_exchange=False)
    mq.connect()

    def agent(i):
        j = i
        while not stop.is_set():
            if mq.xp_queue_size < 64:
                mq.publish_experience(pool[j % len(pool)])
                j += 1
            else:
                time.sleep(0.001)
    threads = [threading.Thread(target=agent, args=(i,), daemon=True) for i in range(agents)]
    for th in threads:
        th.start()
    for it in range(1, 1 + max(1, args.warmup)):
        opt.run_iteration(it)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    env_steps = 0
    for it in range(args.steps):
        m = opt.run_iteration(100 + it)
        env_steps += int(round(m[opt.SPEED_KEY] * m["timing/it"]))
    torch.cuda.synchronize()
    el = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    stop.set()
    if rank == 0:
        sec = float(el.item())
        emit({"metric": "optimizer_steps_per_sec", "value": world * args.steps * epochs / sec, "unit": "steps/s",
                          "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": 1000 * sec / (args.steps * epochs),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "env_steps_per_sec": world * env_steps / sec,
                          "config": {"workload": "BASELINE configs[4]: 40-agent replay stream, in-process broker, reference defaults "
                                                 "(>=1024 seqs x 16 per iteration, 4 epochs, hidden 256 GRU); 'step' = one train() call "
                                                 "including its share of experience prep", "agents": agents * world,
                                     "parallelism": "dp%d" % world}})
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of this run, on the process's real stdout."""
    out = os.fdopen(os.dup(_REAL_STDOUT), "w") if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line: libraries that chat on fd 1 (NCCL prints its version banner there at the first
    # collective) are sent to stderr for the whole run; emit() writes to the saved descriptor.
    global _REAL_STDOUT
    try:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        _REAL_STDOUT = saved
    except OSError:                                  # no usable stderr: keep the plain stdout
        _REAL_STDOUT = None
    args = parse_args()
    cfg = resolve_config(args)
    if cfg.get("stream"):
        run_stream(args, cfg)
        return
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return
    import torch
    import torch.distributed as dist
    from dotaclient_b200 import ops
    from dotaclient_b200.optimizer import DotaOptimizer, ExperienceBatch
    from dotaclient_b200.synthetic import make_rollout, rollout_seed

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl")
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    dev = torch.device("cuda", local_rank)
    B, S, H, cell = cfg["batch"], cfg["seq_len"], cfg["hidden"], cfg["cell"]

    opt = DotaOptimizer(rmq_host="bench", rmq_port=rank, epochs=1, min_seq_per_epoch=B, seq_len=S, learning_rate=5e-5,
                        checkpoint=False, pretrained_model=None, mq_prefetch_count=1, log_dir=tempfile.mkdtemp(),
                        entropy_coef=5e-4, vf_coef=0.5, run_local=True, hidden_size=H, cell=cell)
    # synthetic experience: B rollouts of exactly S steps per rank, prepared by the product's own
    # experiences_from_rollout (old log-probs, values, GAE under the current weights), then stacked once.
    log("optimizer built; preparing %d rollouts of %d steps" % (B, S))
    seqs = []
    with torch.no_grad():
        for i in range(B):
            seqs.extend(opt.experiences_from_rollout(make_rollout(S, rollout_seed(rank, i))))
    batch_dev = ExperienceBatch.from_sequences(seqs, dev)
    del seqs
    batch_host = batch_dev.pin_memory()
    h2d_bytes = batch_host.nbytes()
    torch.cuda.synchronize()
    log("experience batch ready: %.1f MB" % (h2d_bytes / 1e6))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    enqueue = []

    def step_dev():
        opt.train(batch_dev)
        enqueue.append(opt.host_enqueue_s)

    step_e2e_serial = lambda: opt.train(batch_host)  # noqa: E731   (upload, then compute)
    pending = []

    def step_e2e():
        # double-buffered upload through the public API: step k+1's inputs start their H2D copy (from pinned host memory)
        # before step k is launched, so the transfer runs next to step k's kernels; every timed step performs one full upload
        cur = pending.pop() if pending else opt.prefetch(batch_host)
        pending.append(opt.prefetch(batch_host))
        opt.train(cur)

    for _ in range(max(3, args.warmup)):
        step_dev()
    log("warm-up done")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.PROFILE.reset(enabled=True)
    if args.cuda_profiler:
        torch.cuda.profiler.start()
    ms_total = timed(step_dev, args.steps)
    if args.cuda_profiler:
        torch.cuda.profiler.stop()
    log("timed region (HBM-resident) done: %.2f ms/step" % (ms_total / args.steps))
    prof = ops.PROFILE.summary(args.steps)
    prof_bytes = dict(ops.PROFILE.bytes)
    launches = ops.PROFILE.launches
    ops.PROFILE.reset(enabled=False)
    if args.skip_e2e:
        ms_e2e = ms_e2e_serial = float("nan")
    else:
        for _ in range(2):
            step_e2e_serial()
        ms_e2e_serial = timed(step_e2e_serial, args.steps)
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        pending.clear()
    log("timed region (e2e) done: %.2f ms/step" % (ms_e2e / args.steps))
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = ms_total / args.steps
    value = world * 1000.0 / ms_per_step
    e2e_value = world * 1000.0 / (ms_e2e / args.steps)
    tokens = B * S
    G = 4 if cell == "lstm" else 3
    peak, peak_src = measured_peak_hbm()
    # Per-kernel table (CUDA events on the launching stream, averaged per step) with each kernel's ALGORITHMIC HBM bytes
    # (inputs read once + outputs written once) -> achieved GB/s and fraction of the measured HBM peak.
    table = {}
    for name, ms in prof.items():
        nbytes = prof_bytes.get(name, 0) / args.steps
        table[name] = {"ms": ms, "bytes": nbytes, "GBps": (nbytes / (ms * 1e-3) / 1e9) if ms > 0 else 0.0}
        table[name]["frac"] = table[name]["GBps"] / peak
    families = {"tcgen05 3xTF32 GEMM, forward + data gradient (dc_gemm_tf32x3*)": ["gemm_tf32x3"],
                "tcgen05 3xTF32 weight-gradient GEMM (dc_gemm_wgrad_tf32x3*)": ["gemm_wgrad"],
                "recurrence fwd+bwd (dc_rnn_seq_fwd + dc_rnn_seq_bwd)": ["rnn_fwd", "rnn_bwd"]}
    fam = {}
    for label, names in families.items():
        ms = sum(table[n]["ms"] for n in names if n in table)
        by = sum(table[n]["bytes"] for n in names if n in table)
        fam[label] = (ms, by)
    dominant = max(fam, key=lambda k: fam[k][0])                        # the family with the largest share of the step
    dom_ms, dom_bytes = fam[dominant]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None                                                      # measured DRAM bytes per step (ncu), if captured for this shape
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            rec = json.load(f).get("%s_%s" % (args.config, cell))
        if rec and (B, S, H) == (CONFIGS[args.config]["batch"], CONFIGS[args.config]["seq_len"], CONFIGS[args.config]["hidden"]):
            traffic = rec.get(families[dominant][0] if len(families[dominant]) == 1 else "rnn")
    except Exception:
        pass
    rnn_ms, rnn_bytes = fam["recurrence fwd+bwd (dc_rnn_seq_fwd + dc_rnn_seq_bwd)"]
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_step": dom_bytes, "kernel_ms_per_step": dom_ms,
                "share_of_step": dom_ms / ms_per_step, "peak_source": peak_src,
                "recurrence": {"ms_per_step": rnn_ms, "GBps": rnn_bytes / (rnn_ms * 1e-3) / 1e9 if rnn_ms > 0 else 0.0,
                               "note": "latency-bound: 2*S=%d strictly sequential steps per optimizer step" % (2 * S)},
                "kernels": table}
    line = {
        "metric": "optimizer_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(cfg, world, "hbm"),
        "env_steps_per_sec": value * tokens,
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 80,
                "ms_per_step": ms_e2e / args.steps,
                "mode": "double-buffered: DotaOptimizer.prefetch() uploads step k+1 from pinned host memory while step k runs",
                "serial_value": world * 1000.0 / (ms_e2e_serial / args.steps), "serial_ms_per_step": ms_e2e_serial / args.steps},
        "gpu_launches": launches, "host_enqueue_ms_per_step": 1e3 * sum(enqueue[-args.steps:]) / args.steps,
        "roofline": roofline, "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:      # reported baseline: rank 0 at N=1 only
        log("timing the CPU baseline (oracle port)")
        v, cores, sample = cpu_reference_steps_per_sec(cfg, budget_s=20.0)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
