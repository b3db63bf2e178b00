#!/usr/bin/env python
"""bench.py — EM iterations/s of the GMM-EM hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N=1)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N>1)

A "step" is one pass of the EM loop body of the reference (gaussian.cu:532-755):
M-step statistics -> all-reduce of the packed statistics -> host normalisation +
DxD inversions + constants -> parameter upload -> E-step (responsibilities +
log-likelihood), on the synthetic workload of BASELINE.json config 3
(N=10M, D=24, K=64; config 4 shards the same 10M events over N GPUs = strong
scaling).  `value` is measured with the events resident in HBM; `e2e` is measured
through the C ABI from HOST buffers (pinned events H2D, seeding, initial E-step,
the K steps, parameters + log-likelihood D2H all inside the timed region).

`--impl reference` times the reference's algorithm on the host cores (the
sequential-EM CPU port in oracle/, FP32, OpenMP over all cores) on a bounded
slice of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ["NCCL_DEBUG"] = os.environ.get("GMM_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
import __graft_entry__ as entry  # noqa: E402

WORKLOADS = {
    "c3": dict(N=10_000_000, D=24, K=64),      # BASELINE.json configs[2] / [3]
    "c2": dict(N=1_000_000, D=16, K=32),       # configs[1]
    "c1": dict(N=10_000, D=4, K=8),            # configs[0]
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_ready(self, timeout=5.0):
        t_end = time.time() + timeout
        while self.proc and not self.lines and time.time() < t_end:
            time.sleep(0.01)

    def mark(self):
        """Index of the next sample: brackets the timed region (the process is started well before it, so that
        its fork/exec and NVML start-up do not run inside the region)."""
        return len(self.lines)

    def stop(self, begin=0, end=None):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.05)
        self.proc.terminate()
        self.t.join(timeout=2)
        lines = self.lines[begin:(end + 1 if end is not None else None)]
        if not lines:                                   # region shorter than one sampling period: nearest sample
            lines = self.lines[max(0, begin - 1):begin + 1] or self.lines[-1:]
        sm, mx, reasons = [], [], set()
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def run_cpu_baseline(pkg, wl, sample_events, steps=1, warmup=0):
    """Times `steps` EM iterations of the FP32 CPU port on the first
    `sample_events` events; returns (it/s at full N by linear extrapolation, cores, sample text, s/step)."""
    orc = entry.load_oracle("f32")
    N, D, K = wl["N"], wl["D"], wl["K"]
    n = min(sample_events, N)
    ev = pkg.synth.make_blobs(n, D, K, seed=pkg.synth.SEED + 1)
    cl = pkg.Clusters(K, D, n)
    orc.seed(ev, K, cl)
    soa = orc.transpose(ev)
    orc.estep(soa, cl, K)
    for _ in range(warmup):
        orc.mstep(soa, cl, K); orc.constants(cl, K); orc.estep(soa, cl, K)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.mstep(soa, cl, K); orc.constants(cl, K); orc.estep(soa, cl, K)
    dt = (time.perf_counter() - t0) / steps
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    its = 1.0 / (dt * (N / n))
    sample = (f"{steps} EM iteration(s) of the sequential-EM CPU port (oracle/gmm_oracle.c, FP32, OpenMP {cores} threads) "
              f"on {n} of {N} events, D={D}, K={K}; it/s extrapolated linearly to N={N}")
    return its, cores, sample, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GMM_BENCH_WORKLOAD", "c3"), choices=list(WORKLOADS))
    ap.add_argument("--path", default=os.environ.get("GMM_BENCH_PATH", "auto"), choices=["auto", "simt", "tensor"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="events in the cpu_baseline slice (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    metric = "EM iterations/sec at N=10M D=24 K=64; E-step HBM GB/s vs roofline" if args.workload == "c3" else \
        f"EM iterations/sec at N={wl['N']} D={wl['D']} K={wl['K']}"
    pkg = entry.load_package()

    # ---------------- reference arm: CPU port on the host cores ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = args.cpu_sample or 200_000
        its, cores, text, dt = run_cpu_baseline(pkg, wl, sample, steps=max(1, args.steps), warmup=args.warmup)
        line = dict(metric=metric, value=its, unit="it/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=dt * 1e3 * (wl["N"] / min(sample, wl["N"])), higher_is_better=True, scaling=args.scaling,
                    vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                    config=dict(workload=f"{args.workload}: N={wl['N']} D={wl['D']} K={wl['K']} Gaussian blobs (seed {pkg.synth.SEED})"),
                    cpu_baseline=dict(value=its, unit="it/s", cores=cores, kind="port", sample=text),
                    e2e=dict(value=its, unit="it/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    # ---------------- B200 arm --------------------------------------------------
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))
    assert args.gpus == world, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    torch.cuda.set_device(local_rank)
    pkg.load_library()                                   # raises if the CUDA library is missing: no fallback

    N, D, K = wl["N"], wl["D"], wl["K"]
    if args.scaling == "weak":
        N = N * world
    begin, count = pkg.shard_range(N, world, rank)
    # synthetic data: every rank draws the same seeded data set and keeps its shard (pinned host memory)
    ev_all = pkg.synth.make_blobs(N, D, K)
    ev_pinned = torch.empty((count, D), dtype=torch.float32, pin_memory=True)
    ev_pinned.numpy()[...] = ev_all[begin:begin + count]
    del ev_all
    path = {"auto": pkg.PATH_AUTO, "simt": pkg.PATH_SIMT, "tensor": pkg.PATH_TENSOR}[args.path]

    def fresh_nccl_id():
        """A ncclUniqueId can bootstrap ONE communicator: every engine gets its own (rank 0 draws it,
        torch.distributed broadcasts the 128 bytes)."""
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = torch.frombuffer(bytearray(pkg.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(buf, src=0)
        return bytes(buf.numpy().tobytes())

    def make_engine():
        eng = pkg.Engine(None, K, device=local_rank, n_global=N, offset=begin,
                         events_ptr=ev_pinned.data_ptr(), n_local=count, D=D)
        eng.set_option("path", path)
        if world > 1:
            eng.comm_init(world, rank, fresh_nccl_id())
        return eng

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident measurement: value ----
    eng = make_engine()
    seeded = eng.seed(K)
    eng.estep(K)                                         # initial E-step (gaussian.cu:487-523)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                  # before the warm-up: start-up cost stays outside the timed region
    eng.em_iterations(K, args.warmup)
    if rank == 0:
        sampler.wait_ready()
    eng.profile(reset=True)
    barrier()
    m0 = sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    ll = eng.em_iterations(K, args.steps)                # exactly `steps` iterations; returns after a stream sync
    ev1.record()
    torch.cuda.synchronize()
    dt_wall = time.perf_counter() - t0
    dt = ev0.elapsed_time(ev1) * 1e-3                    # device clock around the region (the engine syncs its stream before returning)
    m1 = sampler.mark()
    barrier()
    clocks = sampler.stop(m0, m1) if rank == 0 else None
    dt = max_over_ranks(dt)
    prof = eng.profile()
    value = args.steps / dt

    # ---- end to end from host buffers: e2e ----
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        eng.upload_events(events_ptr=ev_pinned.data_ptr())   # H2D of the pinned shard + device transpose
        eng.set_clusters(K, seeded)                      # parameters H2D
        eng.estep(K)
        ll_e2e = eng.em_iterations(K, args.steps)
        res = eng.get_clusters(K)                        # parameters D2H (log-likelihood already read back)
        torch.cuda.synchronize()
        dt_e2e = max_over_ranks(time.perf_counter() - t0)
        F = 1 + D + D * (D + 1) // 2
        params_bytes = 4 * K * (4 + D + 2 * D * D)
        upload_bytes = 4 * K * (D + D * (D + 1) // 2 + 8)      # packed E-step operand per iteration
        e2e = dict(value=args.steps / dt_e2e, unit="it/s",
                   h2d_bytes_per_step=int((count * D * 4 + params_bytes) / args.steps + upload_bytes),
                   d2h_bytes_per_step=int(8 * (K * F + 1) + params_bytes / args.steps),
                   note=f"one gmm_upload_events (pinned H2D of the {count}x{D} shard) + gmm_set_clusters + gmm_estep + "
                        f"{args.steps} iterations + gmm_get_clusters (D2H) per measurement; per-step bytes amortise the one-time "
                        f"copies; context / NCCL communicator creation is setup and outside the region")
        assert np.isfinite(ll_e2e) and np.all(np.isfinite(res.means[:K]))

    # ---- roofline of the dominant kernels (per-launch CUDA-event times from the engine) ----
    peaks = measured_peaks()
    n_estep = n_mstep = max(1, int(prof["iterations"]))
    estep_ms = prof["estep_ms"] / n_estep
    mstep_ms = prof["mstep_ms"] / n_mstep
    e_bytes = 4.0 * count * (D + K)                       # read X once + write memberships once
    m_flops = 2.0 * count * K * D * D                     # covariance contraction
    e_gbs = e_bytes / (estep_ms * 1e-3) / 1e9 if estep_ms > 0 else 0.0
    m_tfs = m_flops / (mstep_ms * 1e-3) / 1e12 if mstep_ms > 0 else 0.0
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
    # kernels of this repo launched per step: tensor path = mstep_tc_kernel + mstep_tc_finalize_kernel + estep_tc_kernel,
    # SIMT path = mstep_simt_kernel + estep_simt_kernel (cudaMemsetAsync and the NCCL kernel are not counted)
    tensor_m = args.path != "simt" and D in (4, 8, 12, 16, 20, 24)
    launches_per_step = 3 if tensor_m else 2
    # DRAM traffic per launch of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and world == 1:
        tj = json.load(open(tpath))
        key = f"{args.workload}:{'tensor' if args.path != 'simt' else 'simt'}:estep"
        traffic = tj.get(key)
    roofline = dict(kernel="estep", bound="hbm", achieved=e_gbs, peak=peaks["hbm_gbs"], unit="GB/s",
                    frac=e_gbs / peaks["hbm_gbs"], traffic=traffic, peak_source=peaks["source"],
                    ms_per_launch=estep_ms, algorithmic_bytes_per_launch=e_bytes)
    roofline_mstep = dict(kernel="mstep_covariance", bound="tensor", achieved=m_tfs, peak=tf32_peak, unit="TFLOP/s",
                          frac=m_tfs / tf32_peak, peak_source=peaks["source"] + " bf16 sustained / 2 (TF32-equivalent)",
                          ms_per_launch=mstep_ms, algorithmic_flops_per_launch=m_flops)

    if rank != 0:
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if world == 1 and args.cpu_sample > 0:
        its, cores, text, _ = run_cpu_baseline(pkg, wl, args.cpu_sample, steps=1, warmup=0)
        cpu_baseline = dict(value=its, unit="it/s", cores=cores, kind="port", sample=text)

    line = dict(metric=metric, value=value, unit="it/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=dt * 1e3 / args.steps, higher_is_better=True, scaling=args.scaling, vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload=f"{args.workload}: N={N} D={D} K={K} Gaussian blobs (seed {pkg.synth.SEED}), "
                                     f"{count} events on rank 0", path=args.path, l2="inputs (X 0.96 GB + memberships 2.56 GB at c3) exceed the 126 MB L2",
                            parallelism=f"dp{world} (events sharded, one all-reduce of {8 * (K * (1 + D + D * (D + 1) // 2) + 1)} B per step)",
                            arithmetic=("fp32 data and results; tensor path: fp16 hi/lo split operands, fp32 TMEM accumulation, fp64 "
                                        "statistics reduction and host finalisation" if args.path != "simt" else
                                        "fp32 E-step, fp64 M-step statistics and host finalisation")),
                clocks=clocks, e2e=e2e, gpu_launches=launches_per_step * args.steps, roofline=roofline, roofline_mstep=roofline_mstep,
                phases_ms_per_step=dict(estep=estep_ms, mstep=mstep_ms, constants_host=prof["constants_host_ms"] / n_estep,
                                        allreduce=prof["allreduce_ms"] / n_estep, upload=prof["upload_ms"] / n_estep),
                loglik=ll, wall_ms_per_step=dt_wall * 1e3 / args.steps, cpu_baseline=cpu_baseline)
    print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
