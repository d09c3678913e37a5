#!/usr/bin/env python
"""Benchmark of the DNBC hot path (contract in the task prompt / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

A "step" is one Baum-Welch EM iteration (E-step over the rank's resident shard, fp64
all-reduce of the statistics when N > 1, M-step).  The workload is BASELINE.json's scale-out
configuration (N=64 states, T=1000, 16 discrete + 16 continuous variables, 8 Gaussians per
mixture), weak-scaled: each rank holds `--seqs` sequences (default 125,000 = the 1/8 shard,
so 8 GPUs run the configuration's 1,000,000 sequences).
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

from dnbc_scala_b200 import workloads  # noqa: E402

METRIC = "em_iteration_seq_timesteps_per_sec"
UNIT = "seq*timesteps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="scaleout", choices=list(workloads.WORKLOADS))
    ap.add_argument("--seqs", type=int, default=0, help="sequences per GPU (0 = workload S / 8)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    return ap.parse_args()


def config_of(wl, seqs_per_gpu, n_gpus):
    return {"workload": f"{wl.name}: N={wl.N} states, T={wl.T}, D={wl.D} discrete (V={wl.V}) + C={wl.C} continuous, "
                        f"K={wl.K} Gaussians/mixture, {wl.transitions_per_state} transitions/state",
            "sequences_per_gpu": seqs_per_gpu, "sequences_total": seqs_per_gpu * n_gpus, "seq_len": wl.T,
            "parallelism": f"dp{n_gpus} (sequences sharded, fp64 statistics all-reduce per iteration)",
            "l2": "inputs larger than L2 (no flush needed)" if seqs_per_gpu * wl.T * wl.bytes_per_timestep > 126e6
                  else "inputs smaller than L2"}


# ---------------------------------------------------------------------------------
# reference arm / cpu_baseline: the fp64 oracle (a restatement; the JVM original cannot run)
# ---------------------------------------------------------------------------------
def cpu_em_throughput(wl, target_seconds, steps, warmup):
    from oracle import oracle as orc
    orc.build()
    p = workloads.perturbed_params(wl, workloads.true_params(wl))
    m = orc.Model(wl.N, [wl.V] * wl.D, [wl.K] * wl.C)
    for k in ("pi", "A", "B", "w", "mu", "var"):
        getattr(m, k)[...] = p[k].reshape(getattr(m, k).shape)
    cores = orc.num_threads()
    truth = workloads.true_params(wl)

    def draw(S):
        off, disc, cont, labels = workloads.sample_host(wl, truth, S, 4242)
        return orc.Data(off, disc, cont.astype(np.float64), labels)

    probe = draw(max(cores, 8))
    t0 = time.perf_counter()
    st = orc.estep(m, probe)
    dt = time.perf_counter() - t0
    rate = probe.P / dt
    per_step = target_seconds / max(1, steps + warmup)
    S = int(max(cores, min(rate * per_step / wl.T, 20000)))
    data = draw(S)
    times = []
    for i in range(steps + warmup):
        t0 = time.perf_counter()
        st = orc.estep(m, data)
        orc.mstep(m, st, data.S)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    return {"value": data.P / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{S} sequences x T={wl.T} of the same workload ({data.P} timesteps) per step, "
                      f"{steps} timed + {warmup} warm-up EM iterations of the fp64 C/OpenMP oracle "
                      "(restatement: the Scala/Spark reference cannot run here, no JVM)"}, dt


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seqs = args.seqs or max(1, wl.S // 8)
    base, dt = cpu_em_throughput(wl, max(args.cpu_seconds, 10.0) * 2, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wl, seqs, args.gpus), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in ln.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


# ---------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------
def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    from dnbc_scala_b200 import _native as nat
    from dnbc_scala_b200.distributed import StatsAllReduce

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the single JSON line (NCCL prints its version banner there at VERSION/INFO)
        os.environ["NCCL_DEBUG"] = os.environ.get("DNBC_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    seqs = args.seqs or max(1, wl.S // 8)
    truth = workloads.true_params(wl)
    init = workloads.perturbed_params(wl, truth)

    ctx = nat.Context(wl.N, [wl.V] * wl.D, [wl.K] * wl.C, device=local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    ctx.set_params(truth["pi"], truth["A"], truth["B"], truth["w"], truth["mu"], truth["var"])
    ctx.generate(seqs, wl.T, 20260000 + wl.index * 1000 + rank)      # synthetic shard, drawn on the device
    ctx.set_params(init["pi"], init["A"], init["B"], init["w"], init["mu"], init["var"])
    allreduce = StatsAllReduce(ctx)
    total_seqs = seqs * n_gpus
    units_per_step = total_seqs * wl.T

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step():
        ctx.em_estep()
        allreduce()
        ctx.em_mstep(total_seqs)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    # ---- device-resident throughput ("value") -------------------------------------
    ctx.profile_enable(True)
    for _ in range(args.warmup):
        step()
    ctx.profile_read(reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count()
    ms, t0, t1 = timed(step, args.steps, 0)
    launches = ctx.launch_count() - launches0
    fam_ms, fam_n = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    clk = clocks.stop(t0, t1) if rank == 0 else None
    ll_after = float(ctx.em_get_stats()[0])
    value = units_per_step * args.steps / (ms * 1e-3)

    # ---- second BASELINE metric: Viterbi sequences / s (decode of the resident shard, paths
    #      downloaded to the host; upload excluded) ------------------------------------------
    decode = None
    if not args.no_decode:
        # decoded paths and scores land in pinned host buffers (the caller owns the outputs, include/dnbc_b200.h)
        h_path = torch.empty(seqs * wl.T, dtype=torch.int16, pin_memory=True)
        h_score = torch.empty(seqs, dtype=torch.float64, pin_memory=True)

        def decode_step():
            ctx.decode_raw(nat.DECODE_BACKTRACE, nat.PREC_F32, h_path.data_ptr(), h_score.data_ptr(), 0)
        ms_d, _, _ = timed(decode_step, 2, 1)
        assert int(h_path.view(torch.uint8).max()) >= 0 and bool(torch.isfinite(h_score).all())
        decode = {"metric": "viterbi_sequences_per_sec", "value": total_seqs * 2 / (ms_d * 1e-3), "unit": "sequences/s",
                  "mode": "BACKTRACE (true Viterbi, on-device back-trace), f32", "seq_len": wl.T,
                  "ms_per_pass": ms_d / 2, "d2h_bytes_per_pass": int(seqs * wl.T * 2 + seqs * 8)}

    # ---- end to end through the C ABI with host buffers ("e2e") --------------------
    e2e = None
    if not args.no_e2e:
        P = seqs * wl.T
        h_disc = torch.empty((max(wl.D, 1), P), dtype=torch.uint8, pin_memory=True)
        h_cont = torch.empty((max(wl.C, 1), P), dtype=torch.float32, pin_memory=True)
        ctx.download_f32_raw(0, seqs, h_disc.data_ptr() if wl.D else 0, h_cont.data_ptr() if wl.C else 0)
        off = np.arange(seqs + 1, dtype=np.int64) * wl.T
        out = {}

        def e2e_step():
            # upload of chunk k+1 overlaps the kernels of chunk k (dnbc_em_estep_streamed)
            ctx.em_estep_streamed_raw(seqs, off, h_disc.data_ptr() if wl.D else 0, h_cont.data_ptr() if wl.C else 0)
            allreduce()
            ctx.em_mstep(total_seqs)
            out["params"] = ctx.get_params()            # learned parameters back on the host
            out["ll"] = float(ctx.em_get_stats()[0])    # and the iteration's log-likelihood

        ms_e, _, _ = timed(e2e_step, args.steps, min(args.warmup, 1))
        d2h = sum(v.nbytes for v in out["params"].values()) + ctx.stats_size() * 8
        e2e = {"value": units_per_step * args.steps / (ms_e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(P * wl.bytes_per_timestep + off.nbytes), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_e / args.steps}
        del h_disc, h_cont

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu:
        cpu, _ = cpu_em_throughput(wl, args.cpu_seconds, 2, 1)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except OSError:
            pass
        hbm_peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        top = max(fam_ms, key=lambda k: fam_ms[k])
        n_top = max(1, fam_n[top])
        avg_ms = fam_ms[top] / n_top
        pts_per_launch = seqs * wl.T * args.steps / n_top
        achieved = pts_per_launch * wl.bytes_per_timestep / (avg_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                tr = json.load(fh).get(top)
                traffic = tr["bytes_per_point"] * pts_per_launch if tr else None
        except OSError:
            pass
        # the dominant sweeps are bound by the SM's MUFU pipe (ex2 / lg2 / rcp), not by HBM: report that fraction too
        mufu = None
        if top in ("stats_gmm", "emission"):
            sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
            per_clk_sm = wl.N * wl.C * (wl.K + 1) * pts_per_launch / (avg_ms * 1e-3 * sm_mhz * 1e6 * 148)
            mufu = {"kernel": top, "achieved": per_clk_sm, "peak": 15.96, "unit": "MUFU lane-ops/clk/SM",
                    "frac": per_clk_sm / 15.96, "ops_per_timestep": wl.N * wl.C * (wl.K + 1),
                    "peak_source": "measured on this pool's B200 (tools/pipe_peaks.cu, profiles/r01_pipe_peaks.txt)"}
        step_ms = ms / args.steps
        fp32_tflops = wl.flops_per_timestep * seqs * wl.T / (step_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, seqs, n_gpus), "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": which,
                         "avg_launch_ms": avg_ms, "launches_in_timed_region": int(n_top),
                         "share_of_step": fam_ms[top] / max(1e-9, sum(fam_ms.values())),
                         "note": "path is MUFU/FP32-pipe bound (SURVEY 8d), HBM fraction reported for form: see mufu and fp32"},
            "mufu": mufu,
            "fp32": {"algorithmic_tflops_per_gpu": fp32_tflops, "nominal_peak_tflops": 74.4,
                     "frac_of_nominal": fp32_tflops / 74.4, "flops_per_timestep": wl.flops_per_timestep},
            "kernel_ms_per_step": {k: v / args.steps for k, v in fam_ms.items()},
            "loglik_after": ll_after, "decode": decode,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    wl = workloads.WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
