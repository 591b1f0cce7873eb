#!/usr/bin/env python
"""bench.py -- the contract benchmark of the hybrid-ODE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--patients B]

Workload (BASELINE.json configs[1] shape, metric "patient-trajectory solver steps/sec fwd+bwd"):
  2**20 synthetic patients per GPU, dim-8 hybrid RocheODE (generate_data_dim8.py shape: D=8, obs=40, T=15), fixed-step
  RK4 (torchdiffeq 'rk4' = 3/8 rule) with step 1/16 over the full 14-day horizon = 224 solver steps per patient.
  One "step" of the bench = one pass of the hot path over the cohort:
      set_action -> forward solve (with tape) -> fused read-out + masked SSE -> reverse sweep (dL/dy0, dL/dtheta)
      [-> one NCCL all-reduce of the packed parameter gradients when N > 1]
  1 trajectory-step = one RK step of one patient (forward and backward of that step counted once).
  `fwd_only` reports the forward-only sweep of the same cohort (the configs[1] wording) next to the headline.

Data: synthetic, drawn from the reference generator's distributions (dataloader.py:200-266): y0 ~ Exp(scale 0.01),
one dose per patient on a uniform day 0..13 with amount U(0, 10), x ~ N(0,1), mask ~ Bernoulli(0.5); weights: default
nn.Linear init under torch.manual_seed(666), expert scalars at RochConfig defaults.

--impl reference: the reference's CPU path (oracle port of model.py + restated torchdiffeq; /root/reference and the
torchdiffeq package do not exist on the GPU box) on all host cores, each step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1 to every rank.  The host-side work of this file (synthetic cohorts, the CPU reference
# arm) is sized for the host's cores, so give each rank its share BEFORE the OpenMP / MKL runtimes read the variable.
if os.environ.get("OMP_NUM_THREADS") == "1" and "LOCAL_RANK" in os.environ:
    _share = max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))))
    if "reference" in sys.argv:  # --impl reference: rank 0 works alone, the other ranks exit at once
        _share = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(_share)
    os.environ.pop("MKL_NUM_THREADS", None)

# stdout carries exactly one JSON line: NCCL's own messages (its version banner at NCCL_DEBUG=VERSION and above) go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402

D, OBS, T, T_MAX = 8, 40, 15, 14
STEP = 0.0625
N_STEPS = int(T_MAX / STEP)  # 224
METRIC = "patient-trajectory solver steps/sec fwd+bwd"
UNIT = "trajectory-steps/s"
# SURVEY.md 8(d) algorithmic flops (FMA = 2, transcendental = 1): RK4(3/8) step, D = 8: 4*F_f + 18*D with F_f = 104
FLOPS_FWD_STEP = 4 * 104 + 18 * D  # 560
FLOPS_BWD_STEP = 3 * FLOPS_FWD_STEP  # stages recomputed from the tape: fwd + bwd = 4x fwd


def synth_cohort(B, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    y0 = torch.empty(B, D).exponential_(100.0, generator=g)  # scale 0.01
    day = torch.randint(0, T_MAX, (B,), generator=g)
    amt = torch.rand(B, generator=g) * 10.0 + 1e-3
    a = torch.zeros(T, B, 1)
    a[day, torch.arange(B), 0] = amt
    x = torch.randn(T, B, OBS, generator=g)
    mask = (torch.rand(T, B, OBS, generator=g) < 0.5).float()
    out = [y0, a, x, mask]
    if pin:
        out = [t.pin_memory() for t in out]
    if device != "cpu":
        out = [t.to(device) for t in out]
    return out


# -------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def mark(self):
        """Only samples taken after this call count (the sampler is started early: nvidia-smi needs ~1 s to come up)."""
        self.start_index = len(self.lines)

    def __enter__(self):
        self.start_index = 0
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        power = []
        for ln in self.lines[self.start_index:]:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                util = float(f[7])
                if util < 50.0:  # keep samples taken under load only
                    continue
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


def dist_setup(n_gpus):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        import datetime

        # a short collective timeout: a rank-asymmetric bug must fail loudly instead of hanging the box
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    return world, rank, local


# -------------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(sample_patients, repeats, threads=None):
    """Oracle port of the reference path on the host cores: fwd + read-out + masked SSE + backward (autograd)."""
    from oracle import fields as OF

    torch.set_num_threads(threads or os.cpu_count() or 1)
    torch.manual_seed(666)
    dec = OF.OracleDecoder(OBS, D, method="rk4", options={"step_size": STEP})
    y0, a, x, mask = synth_cohort(sample_patients, seed=1)
    times = []
    for _ in range(repeats):
        dec.zero_grad()
        z = y0.clone().requires_grad_(True)
        t0 = time.perf_counter()
        xh, _ = dec(z, a)
        loss = OF.masked_sse(x, xh, mask)
        loss.backward()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return sample_patients * N_STEPS / best, best, torch.get_num_threads()


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fields as OF

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(666)
    B = args.ref_patients
    dec = OF.OracleDecoder(OBS, D, method="rk4", options={"step_size": STEP})
    y0, a, x, mask = synth_cohort(B, seed=1)

    def step():
        dec.zero_grad()
        z = y0.clone().requires_grad_(True)
        xh, _ = dec(z, a)
        loss = OF.masked_sse(x, xh, mask)
        loss.backward()
        return loss.item()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    val = B * N_STEPS * args.steps / el
    sample = "{} patients per step (of the 2^20-patient workload), fwd + read-out/masked-SSE + autograd backward".format(B)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, world=1, note="reference CPU path: oracle port of model.py + restated torchdiffeq 0.2.2 "
                                                   "(neither /root/reference nor torchdiffeq exists on the GPU box)"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(B, world, note=None):
    c = {
        "workload": "dim-8 hybrid RocheODE cohort (generate_data_dim8.py shape), fixed-step RK4 (3/8 rule) h=1/16 over "
                    "14 days = 224 steps/patient; step = set_action + forward solve + fused read-out/masked SSE + "
                    "reverse sweep",
        "patients_per_gpu": B, "latent_dim": D, "obs_dim": OBS, "n_times": T, "solver": "rk4(3/8)", "step_size": STEP,
        "solver_steps_per_patient": N_STEPS, "parallelism": "dp{}".format(world),
        "l2": "inputs larger than L2 (x + mask = {:.2f} GB per GPU, tape {:.2f} GB)".format(
            2 * T * B * OBS * 4 / 1e9, N_STEPS * B * D * 4 / 1e9),
    }
    if note:
        c["note"] = note
    return c


# -------------------------------------------------------------------------------------------------------------------
class StdoutToStderr:
    """stdout must carry exactly ONE JSON line, but NCCL prints its version banner to the C-level stdout when a communicator
    is created.  File descriptor 1 is pointed at stderr for the whole run and restored (after flushing the C stdio buffers)
    just before the line is printed."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def restore(self):
        if self.saved is None:
            return
        sys.stdout.flush()
        try:
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)
        self.saved = None


def run_ours(args):
    import torch.distributed as dist

    quiet = StdoutToStderr()

    import hybrid_ode_neurips_2021_b200 as H
    from hybrid_ode_neurips_2021_b200 import _lib as L
    from hybrid_ode_neurips_2021_b200 import dist as hd

    world, rank, local = dist_setup(args.gpus)
    # torchrun pins OMP_NUM_THREADS=1: give every rank its share of the host cores for the synthetic-cohort generation
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = L.get_lib()  # raises if the CUDA extension is missing: no fallback
    B = args.patients
    torch.manual_seed(666)
    dec = H.RocheExpertDecoder(OBS, D, 1, T_MAX, 1, method="rk4", device=dev,
                               solver_options={"step_size": STEP, "expert_grads": False})
    # host cohort = E2E_CHUNKS mini-batches in pinned memory (what a data loader hands to the training loop, cf.
    # dataloader.py:322 get_split); the device-resident cohort of the kernel-level measurement is their concatenation
    n_chunks = max(1, min(args.e2e_chunks, B // 1024)) if B >= 1024 else 1
    bounds = [(B * i) // n_chunks for i in range(n_chunks + 1)]
    host_chunks = [synth_cohort(bounds[i + 1] - bounds[i], seed=1000 + 97 * rank + i, pin=True) for i in range(n_chunks)]
    y0 = torch.cat([c[0].to(dev) for c in host_chunks], dim=0)
    a, x, mask = (torch.cat([c[k].to(dev) for c in host_chunks], dim=1).contiguous() for k in (1, 2, 3))
    train_params = list(dec.output_function.parameters()) + list(dec.ode.ml_net.parameters())
    B_global = B * world

    def device_step():
        for p in dec.parameters():
            p.grad = None
        z = y0.detach().requires_grad_(True)
        h = dec.solve(z, a)
        loss = H.masked_sse(dec, h, x, mask, n_norm=B_global)
        loss.backward()
        total = hd.allreduce_grads(train_params, extra=loss.detach().reshape(1))
        return loss if total is None else total

    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_step(chunks=None):
        """Public API from HOST buffers: every mini-batch is copied host->device on a copy stream while the previous one
        is solved (forward + loss + backward, gradients accumulate in .grad); one gradient all-reduce; the loss and the
        packed gradients are read back."""
        for p in dec.parameters():
            p.grad = None
        main = torch.cuda.current_stream(dev)
        loss_sum = torch.zeros((), device=dev)
        keep = []
        for (y0_c, a_c, x_c, m_c) in (host_chunks if chunks is None else chunks):
            with torch.cuda.stream(copy_stream):
                dev_c = [t.to(dev, non_blocking=True) for t in (y0_c, a_c, x_c, m_c)]
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            main.wait_event(ready)
            for t in dev_c:
                t.record_stream(main)
            keep.append(dev_c)
            z = dev_c[0].requires_grad_(True)
            h = dec.solve(z, dev_c[1])
            loss = H.masked_sse(dec, h, dev_c[2], dev_c[3], n_norm=B_global)
            loss.backward()
            loss_sum += loss.detach()
        total = hd.allreduce_grads(train_params, extra=loss_sum.reshape(1))
        out = loss_sum if total is None else total
        flat, _ = hd.pack_grads(train_params)
        return float(out.item()), flat.cpu()

    def fwd_only_step():
        with torch.no_grad():
            return dec.solve(y0, a)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, wall = float(tt[0]), float(tt[1]) / 1e3
        return ms, wall

    # every rank runs the SAME sequence of steps (device_step contains a collective when N > 1); only rank 0 samples clocks
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.__enter__()
    for _ in range(max(args.warmup, 3)):
        device_step()
    time.sleep(1.0)  # let nvidia-smi come up before the timed region
    for _ in range(2):
        device_step()
    if sampler is not None:
        sampler.mark()
    ms, wall = timed(device_step, args.steps)  # ms: max over ranks, identical on every rank
    if ms < 1500.0:  # the K timed steps are short: keep the same loop running so that clocks get >= ~1.5 s of samples
        for _ in range(int(1500.0 / max(ms / args.steps, 1e-3)) + 1):
            device_step()
        torch.cuda.synchronize()
    if sampler is not None:
        sampler.__exit__()
    value = B_global * N_STEPS * args.steps / (ms * 1e-3)

    for _ in range(2):
        fwd_only_step()
    ms_f, _ = timed(fwd_only_step, args.steps)
    for _ in range(2):
        e2e_step()
    ms_e, wall_e = timed(e2e_step, args.steps)
    e2e_val = B_global * N_STEPS * args.steps / max(ms_e * 1e-3, wall_e)
    # Secondary: the same end-to-end step when the data loader keeps the 0/1 masks as one byte per entry on the host
    # (masked_sse accepts uint8 / bool masks and widens them on the device; the values, and so the results, are identical).
    # The contract `e2e` above uses the reference's own float32 masks.
    chunks_u8 = [(c[0], c[1], c[2], c[3].to(torch.uint8).pin_memory()) for c in host_chunks]
    for _ in range(2):
        e2e_step(chunks_u8)
    ms_c, wall_c = timed(lambda: e2e_step(chunks_u8), args.steps)
    e2e_compact = {"value": B_global * N_STEPS * args.steps / max(ms_c * 1e-3, wall_c), "unit": UNIT,
                   "ms_per_step": max(ms_c, wall_c * 1e3) / args.steps,
                   "h2d_bytes_per_step": world * sum(t.numel() * t.element_size() for c in chunks_u8 for t in c),
                   "note": "same step with uint8 masks in the pinned host buffers (1 byte instead of 4 per mask entry)"}
    del chunks_u8

    # Every collective of the run is done.  Tear the process group down NOW on every rank: what follows is rank 0's own
    # post-processing (kernel-level timing, CPU baseline), and a rank parked in an NCCL barrier meanwhile would hit the
    # collective timeout (and its spinning host thread slows the CPU baseline down).
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        if rank != 0:
            return

    # ---- per-kernel device times for the roofline (CUDA events on the launching stream, same buffers) -----------------
    roof, extra = None, {}
    if rank == 0:
        from hybrid_ode_neurips_2021_b200 import ops, solver

        dec.ode.set_action(a)
        cfg = ops.make_cfg(L.FIELD_ROCHE, D, L.RK4_38, n_dose=1, expert_grads=False,
                           hill2=solver.hill_exponents_are_two(dec.ode))
        pb = ops.Problem(cfg, 1, B, dec.ode.dosage, dec.ode._dose_t_f32,
                         solver.pack_params(dec.ode, L.FIELD_ROCHE).detach()[None].contiguous(), None)
        tt = torch.arange(0, T_MAX + 1, 1, device=dev, dtype=torch.float32)
        grid = solver.fixed_grid_points(tt.cpu(), STEP).to(dev)
        lin = dec.output_function[0]

        def ev_time(fn, n=7):
            """Median launch duration (CUDA events on the launching stream).  The median, not the mean: the first call
            after the end-to-end phase can include a cudaMalloc of the multi-GB tape between the two events."""
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(n):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); r = fn(); e.record(); torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            return statistics.median(ts), r

        t_fwd, (h, tape) = ev_time(lambda: ops.fixed_fwd(lib, pb, y0, grid, tt, True))
        t_dec, (loss, gh, gw, gb) = ev_time(lambda: ops.decode_sse(lib, h, lin.weight.detach(), lin.bias.detach(), x, mask, B))
        t_bwd, _ = ev_time(lambda: ops.fixed_bwd(lib, pb, grid, tt, gh, tape))
        # tape-free alternative: forward without a tape + the continuous adjoint (odeint_adjoint)
        adj_grid, adj_count = solver.adjoint_grid_points(tt.cpu(), STEP)
        adj_grid, adj_count = adj_grid.to(dev), adj_count.to(dev)
        t_fwd_nt, _ = ev_time(lambda: ops.fixed_fwd(lib, pb, y0, grid, tt, False))
        t_adj, _ = ev_time(lambda: ops.fixed_adjoint(lib, pb, adj_grid, adj_count, h, gh))
        # FP32 FMA peak, measured in this run (MEASURED_PEAKS.json carries HBM and bf16 only)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        out = torch.zeros(1, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        probe = lambda: lib.hode_bench_ffma(sms * 8, 1 << 14, ctypes.c_void_p(out.data_ptr()), stream)  # noqa: E731
        t_probe, flops_probe = ev_time(probe)
        fma_peak = flops_probe / (t_probe * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s"
        bwd_flops = B * N_STEPS * FLOPS_BWD_STEP
        fwd_flops = B * N_STEPS * FLOPS_FWD_STEP
        traffic, traffic_src = ncu_traffic("r01_fixed_bwd_kernel.txt", B)
        roof = {
            "bound": "fp32_fma", "kernel": "fixed_bwd_kernel<Roche<8>, RK4_38> (reverse sweep; largest share of the step)",
            "achieved": bwd_flops / (t_bwd * 1e-3) / 1e12, "peak": fma_peak, "unit": "TFLOP/s",
            "frac": bwd_flops / (t_bwd * 1e-3) / 1e12 / fma_peak, "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": B * 4 * D * (N_STEPS + T + 1),  # tape + grad_h read, grad_y0 written
            "peak_source": "FFMA probe kernel timed in this run ({} SMs; nominal 148 x 128 lanes x 2 x 1.965 GHz = 74.5)".format(sms),
            "algorithmic_flops_per_traj_step": FLOPS_BWD_STEP, "launch_ms": t_bwd,
        }
        dec_bytes = T * B * (2 * OBS + 2 * D) * 4
        extra = {
            "kernels_ms": {"fixed_fwd(+tape)": t_fwd, "decode_sse": t_dec, "fixed_bwd": t_bwd,
                           "fixed_fwd(no tape)": t_fwd_nt, "fixed_adjoint(no tape)": t_adj},
            "adjoint_path": {"note": "odeint_adjoint: forward without a tape + continuous adjoint sweep (4 evals + 4 VJPs "
                                     "per step, same flop count as the reverse sweep); saves the {:.2f} GB tape".format(
                                         B * N_STEPS * D * 4 / 1e9),
                             "value": B * N_STEPS / ((t_fwd_nt + t_dec + t_adj) * 1e-3), "unit": "trajectory-steps/s",
                             "roofline_frac": bwd_flops / (t_adj * 1e-3) / 1e12 / fma_peak},
            "roofline_fwd": {"bound": "fp32_fma", "achieved": fwd_flops / (t_fwd * 1e-3) / 1e12, "peak": fma_peak,
                             "unit": "TFLOP/s", "frac": fwd_flops / (t_fwd * 1e-3) / 1e12 / fma_peak,
                             "algorithmic_flops_per_traj_step": FLOPS_FWD_STEP},
            "roofline_decode_sse": {"bound": "hbm", "achieved": dec_bytes / (t_dec * 1e-3) / 1e9, "peak": hbm_peak,
                                    "unit": "GB/s", "frac": dec_bytes / (t_dec * 1e-3) / 1e9 / hbm_peak,
                                    "peak_source": hbm_src, "algorithmic_bytes_per_launch": dec_bytes},
        }
        del h, tape, gh
        if not args.no_extras and world == 1:  # secondary figures and CPU baselines: N = 1 only
            try:
                extra["other_configs"] = dopri5_extras(lib, dev)
                cpu = dopri5_cpu_baselines()
                for k, v in cpu.items():
                    if k in extra["other_configs"] and "error" not in extra["other_configs"][k]:
                        extra["other_configs"][k]["cpu_baseline"] = v
            except Exception as e:  # secondary figures must never take the headline down
                extra["other_configs"] = {"error": repr(e)}

    if world == 1:
        cpu_val, cpu_s, cores = cpu_reference_rate(args.cpu_patients, 2)
        cpu_baseline = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "{} patients of the same workload, fwd + read-out/masked-SSE + autograd backward, "
                                  "best of 2 ({:.1f} s each)".format(args.cpu_patients, cpu_s)}
    else:  # the CPU baseline is a rank-0, N = 1 measurement (other ranks' host threads would share the cores)
        cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "measured at N=1 only"}
    h2d = world * sum(t.numel() * t.element_size() for c in host_chunks for t in c)  # whole job, like `value`
    n_par = sum(p.numel() for p in train_params)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(B, world),
        "clocks": sampler.summary() if sampler else None,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": world * (4 + 4 * n_par),
                "ms_per_step": max(ms_e, wall_e * 1e3) / args.steps,
                "api": "RocheExpertDecoder.solve + masked_sse + backward over {} pinned host mini-batches, H2D on a copy "
                       "stream overlapped with the previous mini-batch's kernels".format(n_chunks)},
        "gpu_launches": 6 * args.steps,
        "gpu_launches_per_step": {"dose_schedule_kernel": 1, "prep_params_kernel": 2, "fixed_fwd_kernel": 1,
                                  "decode_sse_fast_kernel": 1, "fixed_bwd_kernel": 1},
        "roofline": roof,
        "cpu_baseline": cpu_baseline,
        "fwd_only": {"value": B_global * N_STEPS * args.steps / (ms_f * 1e-3), "unit": UNIT, "ms_per_step": ms_f / args.steps},
        "e2e_uint8_masks": e2e_compact,
    }
    line.update(extra)
    quiet.restore()
    print(json.dumps(line), flush=True)


def ncu_traffic(kernel_file, patients):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), scaled from
    the 2^17-patient capture to this run's cohort (trajectories are independent: traffic is linear in patients)."""
    path = os.path.join(ROOT, "profiles", kernel_file)
    try:
        tot = 0.0
        for ln in open(path):
            f = ln.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[2]]
        if tot <= 0:
            return None, None
        return tot * patients / 131072.0, "profiles/{} (2^17-patient capture x {:.0f})".format(kernel_file, patients / 131072.0)
    except Exception:
        return None, None


def dopri5_extras(lib, dev):
    """Secondary figures (not the headline): adaptive dopri5 at the reference tolerances (rtol 1e-7 / atol 1e-8,
    model.py:1079-1080), forward + tape + reverse sweep, batch-coupled controller = one controller per odeint call.
    C3 shape (run_dim.sh:41): D = 12, groups of 10 patients; C1 shape (sim_config.py:52): D = 6, groups of 50.
    1 trajectory-step = one attempt (accepted or rejected) of one trajectory."""
    import hybrid_ode_neurips_2021_b200 as H
    from hybrid_ode_neurips_2021_b200 import _lib as L
    from hybrid_ode_neurips_2021_b200 import ops, solver

    out = {}
    for name, Dd, groups, batch in (("C3_dim12_dopri5_groups_of_10", 12, 8192, 10), ("C1_dim6_dopri5_groups_of_50", 6, 2048, 50)):
        Bt = groups * batch
        torch.manual_seed(666)
        m = H.RocheODE(Dd, 1, T_MAX, 1, device=dev)
        g = torch.Generator(device=dev).manual_seed(5)
        y0 = torch.empty(Bt, Dd, device=dev).exponential_(100.0, generator=g)
        a = torch.zeros(T, Bt, 1, device=dev)
        a[torch.randint(0, T_MAX, (Bt,), device=dev, generator=g), torch.arange(Bt, device=dev), 0] = \
            torch.rand(Bt, device=dev, generator=g) * 10 + 1e-3
        m.set_action(a)
        tt = torch.arange(0, T_MAX + 1, 1, device=dev, dtype=torch.float64)
        cfg = ops.make_cfg(L.FIELD_ROCHE, Dd, L.DOPRI5, n_dose=1, expert_grads=False, hill2=True, rtol=1e-7, atol=1e-8)
        pb = ops.Problem(cfg, groups, batch, m.dosage, m._dose_t_f32,
                         solver.pack_params(m, L.FIELD_ROCHE).detach()[None].contiguous(), None)

        def ev(fn, n=3):
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(n):
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record(); r = fn(); e_.record(); torch.cuda.synchronize()
                ts.append(s_.elapsed_time(e_))
            return statistics.median(ts), r

        t_f, (h, stats, tape) = ev(lambda: ops.dopri5_fwd(lib, pb, y0, tt, 768))
        st = stats.cpu()
        if int(st[:, 3].max()) != 0:
            out[name] = {"error": "solver status {}".format(st[:, 3].unique().tolist())}
            continue
        gh = torch.randn_like(h)
        t_b, _ = ev(lambda: ops.dopri5_bwd(lib, pb, tt, gh, tape, stats))
        attempts = int((st[:, 0] + st[:, 1]).sum()) * batch
        out[name] = {"value": attempts / ((t_f + t_b) * 1e-3), "unit": UNIT, "fwd_ms": t_f, "bwd_ms": t_b, "patients": Bt,
                     "accepted_per_call": float(st[:, 0].float().mean()), "rejected_per_call": float(st[:, 1].float().mean()),
                     "latent_dim": Dd, "batch_per_odeint_call": batch}
        del h, tape, gh
    return out


def dopri5_cpu_baselines():
    """The reference's CPU path (oracle port) for the two secondary shapes: ONE odeint call of the reference's batch size,
    forward + read-out/masked SSE + autograd backward, trajectory-step attempts counted by the solver itself."""
    from oracle import fields as OF
    from oracle import odeint as OI

    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    for name, Dd, obs, batch in (("C3_dim12_dopri5_groups_of_10", 12, 80, 10), ("C1_dim6_dopri5_groups_of_50", 6, 20, 50)):
        torch.manual_seed(666)
        dec = OF.OracleDecoder(obs, Dd, method="dopri5")
        g = torch.Generator().manual_seed(5)
        y0 = torch.empty(batch, Dd).exponential_(100.0, generator=g)
        a = torch.zeros(T, batch, 1)
        a[torch.randint(0, T_MAX, (batch,), generator=g), torch.arange(batch), 0] = torch.rand(batch, generator=g) * 10 + 1e-3
        x = torch.randn(T, batch, obs, generator=g)
        mask = (torch.rand(T, batch, obs, generator=g) < 0.5).float()
        tr = OI.SolveTrace()
        z = y0.clone().requires_grad_(True)
        t0 = time.perf_counter()
        xh, _ = dec(z, a, trace=tr)
        OF.masked_sse(x, xh, mask).backward()
        el = time.perf_counter() - t0
        att = (tr.accepted + tr.rejected) * batch
        out[name] = {"value": att / el, "unit": UNIT, "seconds": el, "accepted": tr.accepted, "rejected": tr.rejected,
                     "cores": torch.get_num_threads(), "kind": "port", "sample": "one odeint call of {} patients".format(batch)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--patients", type=int, default=1 << 20, help="patients per GPU")
    ap.add_argument("--cpu-patients", type=int, default=8192, help="bounded CPU-baseline sample")
    ap.add_argument("--ref-patients", type=int, default=4096, help="patients per step of --impl reference")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="host mini-batches per step of the end-to-end measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary dopri5 figures")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
