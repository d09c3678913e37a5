#!/usr/bin/env python
"""Benchmark of the DNBC hot path (contract in the task prompt / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

A "step" is one Baum-Welch EM iteration (E-step over the rank's resident shard, fp64
all-reduce of the statistics when N > 1 -- the library's own ncclAllReduce, dnbc_em_allreduce --
and the M-step).  The workload is BASELINE.json's scale-out configuration (N=64 states, T=1000,
16 discrete + 16 continuous variables, 8 Gaussians per mixture), weak-scaled: each rank holds
`--seqs` sequences (default 125,000 = the 1/8 shard, so 8 GPUs run the configuration's
1,000,000 sequences).  Extra keys measured in the same run (seconds of GPU time each):
`full_config` (N=1: the whole 10^6 x T=1000 configuration resident on one GPU), `config3`
(N=10 mixture set, 100,000 sequences STRONG-scaled over the N ranks), `config5` (N=512:
EM on the rank's share of 100,000 sequences, Viterbi on its share of 1,000,000).
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
    ap.add_argument("--no-extras", action="store_true", help="skip the full_config / config3 / config5 legs")
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
def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


def cpu_em_throughput(wl, target_seconds, steps, warmup):
    """The oracle's E+M step on ALL host cores over a bounded sample of the workload.  The thread
    count is set explicitly (torch.distributed.run exports OMP_NUM_THREADS=1) and the sample is a
    whole multiple of it, so the static OpenMP schedule leaves no core idle."""
    from oracle import oracle as orc
    orc.build()
    cores = host_cores()
    orc.set_threads(cores)
    assert orc.num_threads() == cores, (orc.num_threads(), cores)
    p = workloads.perturbed_params(wl, workloads.true_params(wl))
    m = orc.Model(wl.N, [wl.V] * wl.D, [wl.K] * wl.C)
    for k in ("pi", "A", "B", "w", "mu", "var"):
        getattr(m, k)[...] = p[k].reshape(getattr(m, k).shape)
    truth = workloads.true_params(wl)

    def draw(S):
        off, disc, cont, labels = workloads.sample_host(wl, truth, S, 4242)
        return orc.Data(off, disc, cont.astype(np.float64), labels)

    probe = draw(cores)                                  # one sequence per core
    orc.estep(m, probe)                                  # (first call pays page faults / thread start)
    t0 = time.perf_counter()
    orc.estep(m, probe)
    dt = time.perf_counter() - t0
    per_step = target_seconds / max(1, steps + warmup)
    # sequences per core: >= 10 when the budget allows (BASELINE.md asks for a large subsample), never < 2
    per_core = int(min(max(per_step / max(dt, 1e-9), 2), 40))
    if per_core < 10 and 10 * dt * (steps + warmup) <= 4 * target_seconds:
        per_core = 10
    S = per_core * cores
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
            "sample": f"{S} sequences ({per_core} per core) x T={wl.T} of the same workload ({data.P} timesteps) per step, "
                      f"{steps} timed + {warmup} warm-up EM iterations of the fp64 C/OpenMP oracle on {cores} threads "
                      "(restatement: the Scala/Spark reference cannot run here, no JVM)"}, dt


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seqs = args.seqs or max(1, wl.S // 8)
    base, dt = cpu_em_throughput(wl, 150.0, args.steps, args.warmup)
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


def bind_to_gpu_numa_node(local):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of
    the end-to-end leg are allocated node-local (first touch): with 8 ranks pulling 10 GB per step from
    one host the cross-socket hop is what bent the round-1 e2e curve.  Best effort; returns the node."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------
def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    from dnbc_scala_b200 import _native as nat
    from dnbc_scala_b200.distributed import StatsAllReduce, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        # torch.distributed is plumbing only (rendezvous, barriers, max over ranks of the timings); NCCL's own
        # log settings are left as the launcher set them
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    seqs = args.seqs or max(1, wl.S // 8)
    dev = torch.device("cuda", local)

    def make_ctx(w, n_seq, seed):
        truth = workloads.true_params(w)
        init = workloads.perturbed_params(w, truth)
        c = nat.Context(w.N, [w.V] * w.D, [w.K] * w.C, device=local)
        c.set_params(truth["pi"], truth["A"], truth["B"], truth["w"], truth["mu"], truth["var"])
        c.generate(n_seq, w.T, seed)                                   # synthetic shard, drawn on the device
        c.set_params(init["pi"], init["A"], init["B"], init["w"], init["mu"], init["var"])
        return c

    def barrier(c):
        c.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(c, fn, steps, warmup):
        stream = torch.cuda.ExternalStream(c.stream(), device=dev)
        for _ in range(warmup):
            fn()
        barrier(c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier(c)
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    ctx = make_ctx(wl, seqs, 20260000 + wl.index * 1000 + rank)
    allreduce = StatsAllReduce(ctx)            # joins the library-owned communicator (dnbc_comm_init_rank)
    total_seqs = seqs * n_gpus
    units_per_step = total_seqs * wl.T

    def step():
        ctx.em_estep()
        allreduce()                            # dnbc_em_allreduce: ncclAllReduce(sum, f64) on the context's stream
        ctx.em_mstep(total_seqs)

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
    ms, t0, t1 = timed(ctx, step, args.steps, 0)
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
        ms_d, _, _ = timed(ctx, decode_step, 2, 1)
        assert int(h_path.view(torch.uint8).max()) >= 0 and bool(torch.isfinite(h_score).all())
        decode = {"metric": "viterbi_sequences_per_sec", "value": total_seqs * 2 / (ms_d * 1e-3), "unit": "sequences/s",
                  "mode": "BACKTRACE (true Viterbi, on-device back-trace), f32", "seq_len": wl.T,
                  "ms_per_pass": ms_d / 2, "d2h_bytes_per_pass": int(seqs * wl.T * 2 + seqs * 8)}
        del h_path, h_score

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

        ms_e, _, _ = timed(ctx, e2e_step, args.steps, min(args.warmup, 1))
        d2h = sum(v.nbytes for v in out["params"].values()) + ctx.stats_size() * 8
        e2e = {"value": units_per_step * args.steps / (ms_e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(P * wl.bytes_per_timestep + off.nbytes), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_e / args.steps, "host_buffers": "pinned" + (f", NUMA node {numa} (GPU-local)" if numa is not None else "")}
        del h_disc, h_cont
    collectives = None
    if world > 1:
        collectives = {"library": "dnbc_em_allreduce (ncclAllReduce sum f64, library-owned communicator)",
                       "ranks": ctx.comm_info()[0], "bytes": ctx.stats_size() * 8}
    ctx.close()
    del ctx
    try:
        os.sched_setaffinity(0, all_cpus)
    except OSError:
        pass

    # ---- the other BASELINE configurations, seconds each --------------------------------
    extras = {}
    if not args.no_extras:
        def em_leg(w, n_seq, iters, warm, seed):
            c = make_ctx(w, n_seq, seed)
            ar = StatsAllReduce(c)
            tot = int(n_seq)
            if world > 1:
                t = torch.tensor([n_seq], device="cuda", dtype=torch.int64)
                dist.all_reduce(t)
                tot = int(t.item())

            def it():
                c.em_estep()
                ar()
                c.em_mstep(tot)
            l0 = c.launch_count()
            ms_x, _, _ = timed(c, it, iters, warm)
            return c, tot, ms_x / iters, (c.launch_count() - l0) // (iters + warm)

        # config 3: N=10 mixture set, 100,000 sequences split over the ranks (STRONG scaling)
        w3 = workloads.WORKLOADS["gmm100k"]
        lo, hi = shard_range(w3.S, rank, world)
        c3, tot3, ms3, l3 = em_leg(w3, hi - lo, 20, 5, 20263000 + rank)
        c3.close()
        extras["config3"] = {"workload": f"gmm100k: N={w3.N}, T={w3.T}, C={w3.C}, K={w3.K}, {tot3} sequences over {n_gpus} GPU(s)",
                             "scaling": "strong", "ms_per_em_iteration": ms3, "value": tot3 * w3.T / (ms3 * 1e-3), "unit": UNIT,
                             "launches_per_iteration": int(l3),
                             "hbm_floor_ms": tot3 * w3.T * w3.bytes_per_timestep / n_gpus / 6548.2e9 * 1e3}
        # config 5: N=512 dense transitions: EM on the rank's share of 100,000 sequences, Viterbi on its share of 10^6
        w5 = workloads.WORKLOADS["largeN"]
        lo, hi = shard_range(w5.S, rank, world)
        c5, tot5, ms5, l5 = em_leg(w5, hi - lo, 3, 2, 20265000 + rank)
        c5.close()
        n_dec = 1_000_000 // world
        truth5 = workloads.true_params(w5)
        c5 = nat.Context(w5.N, [w5.V] * w5.D, [w5.K] * w5.C, device=local)
        c5.set_params(truth5["pi"], truth5["A"], truth5["B"], truth5["w"], truth5["mu"], truth5["var"])
        c5.generate(n_dec, w5.T, 20265500 + rank)
        h_path = torch.empty(n_dec * w5.T, dtype=torch.int16, pin_memory=True)
        h_score = torch.empty(n_dec, dtype=torch.float64, pin_memory=True)
        ms5d, _, _ = timed(c5, lambda: c5.decode_raw(nat.DECODE_BACKTRACE, nat.PREC_F32, h_path.data_ptr(), h_score.data_ptr(), 0), 1, 1)
        assert bool(torch.isfinite(h_score).all())
        c5.close()
        del h_path, h_score
        extras["config5"] = {"workload": f"largeN: N={w5.N} dense transitions, T={w5.T}, D={w5.D}, C={w5.C}, K={w5.K}",
                             "em": {"sequences_total": tot5, "scaling": "strong", "ms_per_em_iteration": ms5,
                                    "value": tot5 * w5.T / (ms5 * 1e-3), "unit": UNIT, "launches_per_iteration": int(l5),
                                    "gemm_tflops_bf16x3_equiv": 6 * w5.N ** 2 * tot5 * w5.T / n_gpus / (ms5 * 1e-3) / 1e12},
                             "viterbi": {"sequences_total": n_dec * world, "ms_per_pass": ms5d,
                                         "value": n_dec * world / (ms5d * 1e-3), "unit": "sequences/s",
                                         "mode": "BACKTRACE f32, paths and scores copied to pinned host buffers"}}
        # the whole scale-out configuration resident on ONE GPU (no e2e: 80 GB of pinned host memory)
        if n_gpus == 1 and wl.name == "scaleout":
            t_gen = time.perf_counter()
            cf = make_ctx(wl, wl.S, 20264000)
            t_gen = time.perf_counter() - t_gen
            msf, _, _ = timed(cf, lambda: (cf.em_estep(), cf.em_mstep(wl.S)), 2, 1)
            llf = float(cf.em_get_stats()[0])
            cf.close()
            extras["full_config"] = {"workload": f"scaleout at full size: {wl.S} sequences x T={wl.T} resident on 1 GPU "
                                                 f"({wl.S * wl.T * wl.bytes_per_timestep / 1e9:.0f} GB of observations)",
                                     "ms_per_em_iteration": msf / 2, "value": wl.S * wl.T * 2 / (msf * 1e-3), "unit": UNIT,
                                     "generate_s": t_gen, "loglik": llf,
                                     "note": "1 -> 8 GPU strong scaling of the configuration = this time / the 8-GPU ms_per_step"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu:
        cpu, _ = cpu_em_throughput(wl, args.cpu_seconds, 2, 1)

    if rank == 0:
        peaks, pipes = {}, {}
        for name, dst in (("MEASURED_PEAKS.json", peaks), (os.path.join("profiles", "pipe_peaks.json"), pipes)):
            try:
                with open(os.path.join(ROOT, name)) as fh:
                    dst.update(json.load(fh))
            except OSError:
                pass
        hbm_peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        mufu_peak = float(pipes.get("mufu_ops_per_clk_sm", 16.0))
        top = max(fam_ms, key=lambda k: fam_ms[k])
        n_top = max(1, fam_n[top])
        avg_ms = fam_ms[top] / n_top
        pts_per_launch = seqs * wl.T * args.steps / n_top
        hbm_achieved = pts_per_launch * wl.bytes_per_timestep / (avg_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                tr = json.load(fh).get(top)
                traffic = tr["bytes_per_point"] * pts_per_launch if tr else None
        except OSError:
            pass
        step_ms = ms / args.steps
        fp32_tflops = wl.flops_per_timestep * seqs * wl.T / (step_ms * 1e-3) / 1e12
        sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
        share = fam_ms[top] / max(1e-9, sum(fam_ms.values()))
        common = {"kernel": top, "traffic": traffic, "avg_launch_ms": avg_ms, "launches_in_timed_region": int(n_top),
                  "share_of_step": share,
                  "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                          "peak_source": which, "algorithmic_bytes_per_timestep": wl.bytes_per_timestep},
                  "fp32": {"algorithmic_tflops_per_gpu": fp32_tflops, "nominal_peak_tflops": 74.4,
                           "frac_of_nominal": fp32_tflops / 74.4, "flops_per_timestep": wl.flops_per_timestep}}
        if top in ("stats_gmm", "emission"):
            # the binding roof.  The dominant sweeps are bound by the SM's MUFU pipe (ex2 / lg2 / rcp):
            # algorithmic transcendentals N*C*(K+1) per timestep against the measured 15.96 per clock per SM
            ops = wl.N * wl.C * (wl.K + 1)
            per_clk_sm = ops * pts_per_launch / (avg_ms * 1e-3 * sm_mhz * 1e6 * 148)
            roofline = dict(common, bound="mufu", achieved=per_clk_sm, peak=mufu_peak, unit="MUFU lane-ops/clk/SM",
                            frac=per_clk_sm / mufu_peak, ops_per_timestep=ops, sm_mhz_used=sm_mhz,
                            peak_source="measured on this pool's B200 (tools/pipe_peaks.cu, profiles/pipe_peaks.json)")
            # whole step against the same pipe: every algorithmic transcendental of an iteration
            step_ops = wl.mufu_per_timestep
            roofline["step_frac"] = step_ops * seqs * wl.T / (step_ms * 1e-3 * sm_mhz * 1e6 * 148) / mufu_peak
            # what the kernel's bare instruction mix reaches on this GPU (tools/pipe_peaks3.cu: no loads, no bookkeeping)
            body = pipes.get("stats_body_mufu_frac" if top == "stats_gmm" else "emission_body_mufu_frac")
            if body:
                roofline["instruction_mix_ceiling_frac"] = float(body)
        else:
            roofline = dict(common, bound="hbm", achieved=hbm_achieved, peak=hbm_peak, unit="GB/s", frac=hbm_achieved / hbm_peak,
                            peak_source=which)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, seqs, n_gpus), "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline,
            "kernel_ms_per_step": {k: v / args.steps for k, v in fam_ms.items()},
            "loglik_after": ll_after, "decode": decode, "collective": collectives,
            "cpu_baseline": cpu,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
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
