#!/usr/bin/env python
"""bench.py -- the contract benchmark of the hybrid-ODE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--patients B] [--scaling weak|strong]

Workload (BASELINE.json configs[1] shape, metric "patient-trajectory solver steps/sec fwd+bwd"):
  2**20 synthetic patients per GPU, dim-8 hybrid RocheODE (generate_data_dim8.py shape: D=8, obs=40, T=15), fixed-step
  RK4 (torchdiffeq 'rk4' = 3/8 rule) with step 1/16 over the full 14-day horizon = 224 solver steps per patient.
  One "step" of the bench = one pass of the hot path over the cohort:
      set_action -> forward solve with tape, read-out + masked SSE consumed at the output times (one launch)
                 -> reverse sweep (dL/dy0, dL/dtheta)
      [-> one NCCL all-reduce of the packed parameter gradients when N > 1]
  1 trajectory-step = one RK step of one patient (forward and backward of that step counted once).
  `fwd_only` reports the forward-only sweep of the same cohort (the configs[1] wording) next to the headline.

Data: synthetic, drawn from the reference generator's distributions (dataloader.py:200-266): y0 ~ Exp(scale 0.01),
one dose per patient on a uniform day 0..13 with amount U(0, 10), x ~ N(0,1), mask ~ Bernoulli(0.5); weights: default
nn.Linear init under torch.manual_seed(666), expert scalars at RochConfig defaults.

End to end (`e2e`): the same step through the public API from pinned HOST buffers -- float32 y0 / actions / measurements and
the 0/1 masks as ONE BYTE per entry (the reference's masks are float32 0/1 tensors, dataloader.py:264-266; masked_sse takes
uint8 / bool masks and the results are identical) -- H2D copies on a copy stream inside the timed region, loss + packed
gradients read back.  `e2e_float_masks` is the same with 4-byte masks, `e2e_resident` the reference's actual usage
(dg.set_device(device), run_simulation.py:65: the cohort is moved to the device once per run, not once per step).

--scaling strong: `--patients` is the TOTAL cohort, split over the ranks (default weak: `--patients` per GPU).

--impl reference: the reference's CPU path on all host cores, each step a bounded sample of the same workload: the
reference's own model.py classes (RocheODE vector field + set_action + output_function, from the byte-identical copy in
baseline/_ref made by oracle/install_reference.py) driven by the restated torchdiffeq 0.2.2 (oracle/odeint.py; the package
is not installable offline) -- kind "reference"; the oracle port of model.py if baseline/_ref is absent -- kind "port".
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
def reference_decoder():
    """The reference's CPU implementation of the path for the C2 workload.  Returns (kind, decoder-like callable, loss).

    kind "reference": model.RocheODE / output_function of the reference's own model.py (baseline/_ref or /root/reference),
    integrated by the restated torchdiffeq with options={'step_size': 1/16}.  (RocheExpertDecoder.forward itself never
    forwards a step size -- model.py:1116-1118 -- and the default grid h = 1 diverges to NaN, SURVEY.md fact 6, so the
    decoder's three statements are issued here with the option added.)  kind "port": the oracle's restatement of model.py."""
    from oracle import fields as OF
    from oracle import odeint as OI
    from oracle import refload

    torch.manual_seed(666)
    if refload.available():
        M = refload.load("model")
        dec = M.RocheExpertDecoder(OBS, D, 1, T_MAX, 1, roche=True, method="rk4", device=torch.device("cpu"))

        def run(z, a):
            dec.ode.set_action(a)  # the reference's O(B) Python loop, model.py:495-507
            h = OI.odeint(dec.ode, z, dec.t, rtol=1e-7, atol=1e-8, method="rk4", options={"step_size": STEP})
            return dec.output_function(h), h

        return "reference", dec, run, (lambda x, xh, m: torch.sum((x - xh) ** 2 * m) / x.shape[1])
    dec = OF.OracleDecoder(OBS, D, method="rk4", options={"step_size": STEP})
    return "port", dec, (lambda z, a: dec(z, a)), OF.masked_sse


def cpu_reference_rate(sample_patients, repeats, threads=None):
    """The reference path on the host cores: set_action + fwd + read-out + masked SSE + backward (autograd)."""
    torch.set_num_threads(threads or os.cpu_count() or 1)
    kind, dec, run, lossf = reference_decoder()
    y0, a, x, mask = synth_cohort(sample_patients, seed=1)
    times = []
    for _ in range(repeats):
        dec.zero_grad()
        z = y0.clone().requires_grad_(True)
        t0 = time.perf_counter()
        xh, _ = run(z, a)
        loss = lossf(x, xh, mask)
        loss.backward()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return sample_patients * N_STEPS / best, best, torch.get_num_threads(), kind


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    B = args.ref_patients
    kind, dec, run, lossf = reference_decoder()
    y0, a, x, mask = synth_cohort(B, seed=1)

    def step():
        dec.zero_grad()
        z = y0.clone().requires_grad_(True)
        xh, _ = run(z, a)
        loss = lossf(x, xh, mask)
        loss.backward()
        return loss.item()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    val = B * N_STEPS * args.steps / el
    what = ("reference model.py classes (baseline/_ref) + restated torchdiffeq 0.2.2" if kind == "reference"
            else "oracle port of model.py + restated torchdiffeq 0.2.2 (baseline/_ref absent)")
    sample = "{} patients per step (of the 2^20-patient workload), set_action + fwd + read-out/masked-SSE + autograd backward".format(B)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, world=1, note="reference CPU path: " + what),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
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


def numa_local_affinity(dev_index):
    """Pin this process to the cores of the GPU's NUMA node BEFORE any host buffer is allocated or pinned (first-touch page
    placement): with all ranks on one node, 8 GPUs' worth of pinned H2D traffic crosses one socket's memory controllers
    (round 1: 182 GB/s aggregate at N = 8).  Returns a description for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        bus = "{:04x}:{:02x}:{:02x}.0".format(pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/{}/numa_node".format(bus)).read().strip())
        if node < 0:
            return {"pci": bus, "numa_node": node, "pinned": False}
        cpus = set()
        for part in open("/sys/devices/system/node/node{}/cpulist".format(node)).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0) or cpus
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"pci": bus, "numa_node": node, "pinned": bool(cpus), "cores": len(cpus)}
    except Exception as e:  # best effort: containers may hide sysfs
        return {"error": repr(e), "pinned": False}


def run_ours(args):
    import torch.distributed as dist

    quiet = StdoutToStderr()

    import hybrid_ode_neurips_2021_b200 as H
    from hybrid_ode_neurips_2021_b200 import _lib as L
    from hybrid_ode_neurips_2021_b200 import dist as hd

    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = numa_local_affinity(local)
    # torchrun pins OMP_NUM_THREADS=1: give every rank its share of the host cores for the synthetic-cohort generation
    torch.set_num_threads(max(1, min(len(os.sched_getaffinity(0)), (os.cpu_count() or 1) // max(world, 1))))
    lib = L.get_lib()  # raises if the CUDA extension is missing: no fallback
    strong = args.scaling == "strong"
    B = args.patients // world if strong else args.patients  # patients per GPU
    B_global = B * world
    torch.manual_seed(666)
    dec = H.RocheExpertDecoder(OBS, D, 1, T_MAX, 1, method="rk4", device=dev,
                               solver_options={"step_size": STEP, "expert_grads": False})
    for n_, p_ in dec.ode.named_parameters():  # the optimizer of the reference never receives the 13 expert scalars
        if not n_.startswith("ml_net"):        # (run_simulation.py:125-129): trained are ml_net + output_function
            p_.requires_grad_(False)
    train_params = list(dec.output_function.parameters()) + list(dec.ode.ml_net.parameters())
    fg = hd.FlatGrads(train_params, extra=1)  # gradients live in ONE flat buffer: no pack / unpack kernels, one collective

    # ---- multi-GPU correctness: the all-reduced gradient of a cohort split N ways == the 1-GPU gradient of the cohort ------
    grad_check = None
    if world > 1:
        Bc = 1 << 16
        y0c, ac, xc, mc = [t.to(dev) for t in synth_cohort(Bc, seed=4242)]

        def cohort_grad(lo, hi):
            fg.zero_()
            z = y0c[lo:hi].detach().requires_grad_(True)
            loss = dec.loss(z, ac[:, lo:hi].contiguous(), xc[:, lo:hi].contiguous(), mc[:, lo:hi].contiguous(), n_norm=Bc)
            loss.backward()
            fg.extra.copy_(loss.detach().reshape(1))

        cohort_grad(0, Bc)
        full = fg.flat.clone()
        lo, hi = hd.shard_range(Bc, rank, world)
        cohort_grad(lo, hi)
        fg.allreduce()
        err = float((fg.grads - full[:-1]).abs().max() / full[:-1].abs().max())
        lerr = float((fg.extra - full[-1:]).abs().max() / full[-1:].abs().max())
        grad_check = {"patients": Bc, "ranks": world, "grad_relerr_inf": err, "loss_relerr": lerr, "tol": 5e-4,
                      "ok": bool(err <= 5e-4 and lerr <= 1e-5)}
        if not grad_check["ok"]:
            raise AssertionError("all-reduced sharded gradient differs from the single-GPU gradient: {}".format(grad_check))
        del y0c, ac, xc, mc, full

    # host cohort = E2E_CHUNKS mini-batches in pinned memory (what a data loader hands to the training loop, cf.
    # dataloader.py:322 get_split); the device-resident cohort of the kernel-level measurement is their concatenation
    n_chunks = max(1, min(args.e2e_chunks, B // 1024)) if B >= 1024 else 1
    bounds = [(B * i) // n_chunks for i in range(n_chunks + 1)]
    host_f32 = [synth_cohort(bounds[i + 1] - bounds[i], seed=1000 + 97 * rank + i, pin=True) for i in range(n_chunks)]
    host_u8 = [(c[0], c[1], c[2], c[3].to(torch.uint8).pin_memory()) for c in host_f32]  # contract path: 1-byte masks
    y0 = torch.cat([c[0].to(dev) for c in host_f32], dim=0)
    a, x, mask = (torch.cat([c[k].to(dev) for c in host_f32], dim=1).contiguous() for k in (1, 2, 3))

    def device_step():
        fg.zero_()
        z = y0.detach().requires_grad_(True)
        loss = dec.loss(z, a, x, mask, n_norm=B_global)  # solve + read-out + masked SSE: one forward launch
        loss.backward()
        fg.extra.copy_(loss.detach().reshape(1))
        return fg.allreduce()

    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_step(chunks):
        """Public API from HOST buffers: every mini-batch is copied host->device on a copy stream while the previous one
        is solved (forward + loss + backward, gradients accumulate in the flat buffer); one gradient all-reduce; the loss
        and the packed gradients are read back."""
        fg.zero_()
        main = torch.cuda.current_stream(dev)
        # all copies are queued on the copy stream first (one event per mini-batch): set_action's host read of the dose count
        # would otherwise hold back the NEXT mini-batch's copies until the current one has landed, leaving the copy engine idle
        # between mini-batches
        staged = []
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_stream(main)
            for (y0_c, a_c, x_c, m_c) in chunks:
                dev_c = [t.to(dev, non_blocking=True) for t in (y0_c, a_c, x_c, m_c)]
                ready = torch.cuda.Event()
                ready.record(copy_stream)
                staged.append((dev_c, ready))
        for dev_c, ready in staged:
            main.wait_event(ready)
            for t in dev_c:
                t.record_stream(main)
            z = dev_c[0].requires_grad_(True)
            loss = dec.loss(z, dev_c[1], dev_c[2], dev_c[3], n_norm=B_global)
            loss.backward()
            fg.extra.add_(loss.detach().reshape(1))
        fg.allreduce()
        return fg.flat.cpu()  # loss + packed gradients: the device->host read of the step's result

    def resident_step():
        """The reference's actual usage: the cohort lives on the device (dg.set_device, run_simulation.py:65); per step only
        the loss and the gradients come back."""
        device_step()
        return fg.flat.cpu()

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

    rate = lambda ms, wall=0.0: B_global * N_STEPS * args.steps / max(ms * 1e-3, wall)  # noqa: E731

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
    value = rate(ms)

    for _ in range(2):
        fwd_only_step()
    ms_f, _ = timed(fwd_only_step, args.steps)
    for _ in range(2):
        e2e_step(host_u8)
    ms_e, wall_e = timed(lambda: e2e_step(host_u8), args.steps)
    for _ in range(2):
        e2e_step(host_f32)
    ms_c, wall_c = timed(lambda: e2e_step(host_f32), args.steps)
    for _ in range(2):
        resident_step()
    ms_r, wall_r = timed(resident_step, args.steps)
    bytes_of = lambda chunks: world * sum(t.numel() * t.element_size() for c in chunks for t in c)  # noqa: E731
    n_par = fg.n_params
    e2e = {"value": rate(ms_e, wall_e), "unit": UNIT, "h2d_bytes_per_step": bytes_of(host_u8),
           "d2h_bytes_per_step": world * 4 * (n_par + 1), "ms_per_step": max(ms_e, wall_e * 1e3) / args.steps,
           "api": "RocheExpertDecoder.solve + masked_sse + backward over {} pinned host mini-batches per GPU (float32 y0 / "
                  "actions / measurements, uint8 masks), H2D on a copy stream overlapped with the previous mini-batch's "
                  "kernels; loss + packed gradients read back".format(n_chunks)}
    e2e_float = {"value": rate(ms_c, wall_c), "unit": UNIT, "ms_per_step": max(ms_c, wall_c * 1e3) / args.steps,
                 "h2d_bytes_per_step": bytes_of(host_f32), "note": "same step with the reference's float32 0/1 masks on the host"}
    e2e_resident = {"value": rate(ms_r, wall_r), "unit": UNIT, "ms_per_step": max(ms_r, wall_r * 1e3) / args.steps,
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": world * 4 * (n_par + 1),
                    "note": "cohort resident on the device (the reference's dg.set_device usage, run_simulation.py:65); per "
                            "step only the loss + packed gradients are read back"}

    # ---- host -> device ceiling of this box: every rank copies its own pinned 1 GiB buffer, all ranks at once ---------------
    probe = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    for _ in range(2):
        dst.copy_(probe, non_blocking=True)
    ms_p, _ = timed(lambda: dst.copy_(probe, non_blocking=True), 5)
    h2d_probe = {"gbs_per_gpu_all_ranks_concurrently": 5 * (1 << 30) / (ms_p * 1e-3) / 1e9, "ranks": world,
                 "aggregate_gbs": world * 5 * (1 << 30) / (ms_p * 1e-3) / 1e9, "numa": numa,
                 "note": "cudaMemcpyAsync of one pinned 1 GiB buffer per rank, max over ranks of the time"}
    del probe, dst

    # Every collective of the run is done.  Tear the process group down NOW on every rank: what follows is rank 0's own
    # post-processing (kernel-level timing, CPU baseline), and a rank parked in an NCCL barrier meanwhile would hit the
    # collective timeout (and its spinning host thread slows the CPU baseline down).
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        if rank != 0:
            return
    del host_u8

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
        # the training step's forward launch: solve + read-out + masked SSE consumed at the output times (no h, no decode pass)
        t_fused, _ = ev_time(lambda: ops.fixed_fwd_sse(lib, pb, y0, grid, tt, lin.weight.detach(), lin.bias.detach(), x, mask, B,
                                                       want_tape=True))
        # tape-free alternative: forward without a tape + the continuous adjoint (odeint_adjoint)
        adj_grid, adj_count = solver.adjoint_grid_points(tt.cpu(), STEP)
        adj_grid, adj_count = adj_grid.to(dev), adj_count.to(dev)
        t_fwd_nt, _ = ev_time(lambda: ops.fixed_fwd(lib, pb, y0, grid, tt, False))
        t_adj, _ = ev_time(lambda: ops.fixed_adjoint(lib, pb, adj_grid, adj_count, h, gh))
        fma_peak, peak_src = ffma_peak(lib, dev, ev_time)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s"
        bwd_flops = B * N_STEPS * FLOPS_BWD_STEP
        fwd_flops = B * N_STEPS * FLOPS_FWD_STEP
        traffic, traffic_src = ncu_traffic("r02_fixed_bwd_kernel.txt", B)
        roof = {
            "bound": "fp32_fma", "kernel": "fixed_bwd_kernel<Roche<8>, RK4_38> (reverse sweep; largest share of the step)",
            "achieved": bwd_flops / (t_bwd * 1e-3) / 1e12, "peak": fma_peak, "unit": "TFLOP/s",
            "frac": bwd_flops / (t_bwd * 1e-3) / 1e12 / fma_peak, "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": B * 4 * D * (N_STEPS + T + 1),  # tape + grad_h read, grad_y0 written
            "peak_source": peak_src,
            "algorithmic_flops_per_traj_step": FLOPS_BWD_STEP, "launch_ms": t_bwd,
        }
        dec_bytes = T * B * (2 * OBS + 2 * D) * 4
        extra = {
            "kernels_ms": {"fixed_fwd_sse(+tape) [in the step]": t_fused, "fixed_bwd [in the step]": t_bwd,
                           "fixed_fwd(+tape) [two-launch path]": t_fwd, "decode_sse [two-launch path]": t_dec,
                           "fixed_fwd(no tape)": t_fwd_nt, "fixed_adjoint(no tape)": t_adj},
            "adjoint_path": {"note": "odeint_adjoint: forward without a tape + continuous adjoint sweep (4 evals + 4 VJPs "
                                     "per step, same flop count as the reverse sweep); saves the {:.2f} GB tape".format(
                                         B * N_STEPS * D * 4 / 1e9),
                             "value": B * N_STEPS / ((t_fwd_nt + t_dec + t_adj) * 1e-3), "unit": "trajectory-steps/s",
                             "roofline_frac": bwd_flops / (t_adj * 1e-3) / 1e12 / fma_peak},
            "roofline_fwd_fused": {"bound": "fp32_fma", "kernel": "fixed_fwd_sse_kernel<Roche<8>, RK4_38>",
                                   "achieved": (fwd_flops + B * T * (6 * OBS * D + 5 * OBS)) / (t_fused * 1e-3) / 1e12, "peak": fma_peak,
                                   "unit": "TFLOP/s", "frac": (fwd_flops + B * T * (6 * OBS * D + 5 * OBS)) / (t_fused * 1e-3) / 1e12 / fma_peak,
                                   "launch_ms": t_fused, "hbm_gbs": (dec_bytes - B * T * D * 4 + B * N_STEPS * D * 4) / (t_fused * 1e-3) / 1e9,
                                   "note": "flops = 560 per trajectory-step + (6 obs D + 5 obs) per (output time, trajectory) for "
                                           "read-out, grad_h and grad_W; streams x + mask in, tape + grad_h out"},
            "roofline_fwd": {"bound": "fp32_fma", "achieved": fwd_flops / (t_fwd * 1e-3) / 1e12, "peak": fma_peak,
                             "unit": "TFLOP/s", "frac": fwd_flops / (t_fwd * 1e-3) / 1e12 / fma_peak,
                             "algorithmic_flops_per_traj_step": FLOPS_FWD_STEP},
            "roofline_decode_sse": {"bound": "hbm", "achieved": dec_bytes / (t_dec * 1e-3) / 1e9, "peak": hbm_peak,
                                    "unit": "GB/s", "frac": dec_bytes / (t_dec * 1e-3) / 1e9 / hbm_peak,
                                    "peak_source": hbm_src, "algorithmic_bytes_per_launch": dec_bytes},
        }
        del h, tape, gh
        if not args.no_extras and world == 1:  # secondary figures and CPU baselines: N = 1 only
            del y0, a, x, mask
            torch.cuda.empty_cache()
            try:
                oc = dopri5_extras(lib, dev, fma_peak)
                cpu = dopri5_cpu_baselines()
                for k, v in cpu.items():
                    if k in oc and "error" not in oc[k]:
                        oc[k]["cpu_baseline"] = v
                extra["other_configs"] = oc
                for key, name in (("roofline_c3", "C3_dim12_dopri5_groups_of_10"), ("roofline_c1", "C1_dim6_dopri5_groups_of_50")):
                    if name in oc and "roofline" in oc[name]:
                        extra[key] = oc[name]["roofline"]
                extra["other_configs"]["C4_ensemble_members"] = ensemble_extra(dev)
                torch.cuda.empty_cache()
                extra["other_configs"].update(c5_extras(dev, fma_peak))
            except Exception as e:  # secondary figures must never take the headline down
                extra.setdefault("other_configs", {})["error"] = repr(e)

    if world == 1:
        cpu_val, cpu_s, cores, kind = cpu_reference_rate(args.cpu_patients, 2)
        cpu_baseline = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": "{} patients of the same workload, set_action + fwd + read-out/masked-SSE + autograd backward, "
                                  "best of 2 ({:.1f} s each)".format(args.cpu_patients, cpu_s)}
    else:  # the CPU baseline is a rank-0, N = 1 measurement (other ranks' host threads would share the cores)
        cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "measured at N=1 only"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(B, world),
        "clocks": sampler.summary() if sampler else None,
        "e2e": e2e,
        "gpu_launches": 6 * args.steps,
        "gpu_launches_per_step": {"dose_schedule_kernel": 1, "prep_params_kernel": 2, "prep_readout_kernel": 1,
                                  "fixed_fwd_sse_kernel": 1, "fixed_bwd_kernel": 1},
        "roofline": roof,
        "cpu_baseline": cpu_baseline,
        "fwd_only": {"value": rate(ms_f), "unit": UNIT, "ms_per_step": ms_f / args.steps},
        "e2e_float_masks": e2e_float,
        "e2e_resident": e2e_resident,
        "h2d_probe": h2d_probe,
        "grad_check": grad_check,
    }
    line.update(extra)
    quiet.restore()
    print(json.dumps(line), flush=True)


def ffma_peak(lib, dev, ev_time):
    """FP32 FMA peak measured in this run (MEASURED_PEAKS.json carries HBM and bf16 only)."""
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    out = torch.zeros(1, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    probe = lambda: lib.hode_bench_ffma(sms * 8, 1 << 14, ctypes.c_void_p(out.data_ptr()), stream)  # noqa: E731
    t_probe, flops_probe = ev_time(probe)
    return flops_probe / (t_probe * 1e-3) / 1e12, ("FFMA probe kernel timed in this run ({} SMs; nominal 148 x 128 lanes x 2 x "
                                                   "1.965 GHz = 74.5)".format(sms))


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


def dopri5_extras(lib, dev, fma_peak):
    """Secondary configurations (not the headline): adaptive dopri5 at the reference tolerances (rtol 1e-7 / atol 1e-8,
    model.py:1079-1080), forward + tape + reverse sweep, batch-coupled controller = one controller per odeint call.
    C3 shape (BASELINE configs[2], run_dim.sh:41): D = 12, odeint calls of 10 patients; C1 shape (sim_config.py:52): D = 6,
    calls of 50.  1 trajectory-step = one attempt (accepted or rejected) of one trajectory.
    Roofline flops: attempt 6 F_f + 64 D, + 44 D per accepted step; reverse sweep 3 x the accepted-step flops."""
    import hybrid_ode_neurips_2021_b200 as H
    from hybrid_ode_neurips_2021_b200 import _lib as L
    from hybrid_ode_neurips_2021_b200 import ops, solver

    out = {}
    for name, Dd, groups, batch in (("C3_dim12_dopri5_groups_of_10", 12, 65536, 10), ("C1_dim6_dopri5_groups_of_50", 6, 8192, 50)):
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

        t_f, (h, stats, tape) = ev(lambda: ops.dopri5_fwd(lib, pb, y0, tt, 448))
        st = stats.cpu()
        if int(st[:, 3].max()) != 0:
            out[name] = {"error": "solver status {}".format(st[:, 3].unique().tolist())}
            continue
        gh = torch.randn_like(h)
        t_b, _ = ev(lambda: ops.dopri5_bwd(lib, pb, tt, gh, tape, stats))
        acc, rej = int(st[:, 0].sum()) * batch, int(st[:, 1].sum()) * batch
        Ff = 29 + 2 * Dd * (Dd - 4) + (Dd - 4) + 3 + (Dd - 4)
        fl_att, fl_acc = 6 * Ff + 64 * Dd, 44 * Dd
        fwd_tf = ((acc + rej) * fl_att + acc * fl_acc) / (t_f * 1e-3) / 1e12
        bwd_tf = 3 * acc * (fl_att + fl_acc) / (t_b * 1e-3) / 1e12
        out[name] = {"value": (acc + rej) / ((t_f + t_b) * 1e-3), "unit": UNIT, "fwd_ms": t_f, "bwd_ms": t_b, "patients": Bt,
                     "odeint_calls": groups, "accepted_per_call": float(st[:, 0].float().mean()),
                     "rejected_per_call": float(st[:, 1].float().mean()), "latent_dim": Dd, "batch_per_odeint_call": batch,
                     "roofline": {"bound": "fp32_fma", "peak": fma_peak, "unit": "TFLOP/s",
                                  "fwd": {"kernel": "dopri5_fwd_seg_kernel" if batch <= 16 else "dopri5_fwd_kernel",
                                          "achieved": fwd_tf, "frac": fwd_tf / fma_peak, "launch_ms": t_f,
                                          "flops_per_attempt": fl_att, "flops_per_accepted_step_extra": fl_acc},
                                  "bwd": {"kernel": "dopri5_bwd_kernel", "achieved": bwd_tf, "frac": bwd_tf / fma_peak,
                                          "launch_ms": t_b, "flops_per_accepted_step": 3 * (fl_att + fl_acc)},
                                  "frac": min(fwd_tf, bwd_tf) / fma_peak}}
        # tape-free alternative: forward without a tape + the adaptive continuous adjoint (odeint_adjoint, 'seminorm'), on the
        # first 8 192 odeint calls (one CTA per call: no lane-segment variant of this kernel yet)
        try:
            ga = min(groups, 8192)
            acfg = ops.make_cfg(L.FIELD_ROCHE, Dd, L.DOPRI5, n_dose=1, expert_grads=False, hill2=True, rtol=1e-7, atol=1e-8,
                                adj_seminorm=True)
            apb = ops.Problem(acfg, ga, batch, m.dosage[:ga * batch], m._dose_t_f32[:ga * batch], pb.params, None)
            t_f0, (h0, st0, _) = ev(lambda: ops.dopri5_fwd(lib, apb, y0[:ga * batch], tt, 0))
            gh0 = gh[:, :ga * batch].contiguous()
            t_a, (_, _, sta) = ev(lambda: ops.dopri5_adjoint(lib, apb, tt, h0, gh0))
            sta = sta.cpu()
            out[name]["adjoint_path"] = {
                "odeint_calls": ga, "fwd_no_tape_ms": t_f0, "adjoint_ms": t_a, "status_ok": bool(int(sta[:, 3].max()) == 0),
                "adjoint_accepted_per_call": float(sta[:, 0].float().mean()), "adjoint_rejected_per_call": float(sta[:, 1].float().mean()),
                "value": int((st0.cpu()[:, 0] + st0.cpu()[:, 1]).sum()) * batch / ((t_f0 + t_a) * 1e-3), "unit": UNIT,
                "note": "odeint_adjoint(method='dopri5', adjoint_options={'norm': 'seminorm'}): no tape is allocated"}
            if Dd <= 8:  # torchdiffeq's default mixed norm (built for RocheODE up to latent_dim 8, batch-coupled controller)
                mcfg = ops.make_cfg(L.FIELD_ROCHE, Dd, L.DOPRI5, n_dose=1, expert_grads=False, hill2=True, rtol=1e-7, atol=1e-8)
                mpb = ops.Problem(mcfg, ga, batch, m.dosage[:ga * batch], m._dose_t_f32[:ga * batch], pb.params, None)
                t_m, (_, _, stm) = ev(lambda: ops.dopri5_adjoint(lib, mpb, tt, h0, gh0))
                stm = stm.cpu()
                out[name]["adjoint_path"]["mixed_norm"] = {
                    "adjoint_ms": t_m, "status_ok": bool(int(stm[:, 3].max()) == 0),
                    "adjoint_accepted_per_call": float(stm[:, 0].float().mean()),
                    "adjoint_rejected_per_call": float(stm[:, 1].float().mean()),
                    "note": "torchdiffeq's default adjoint norm: the parameter adjoints take part in the error control"}
            del h0, gh0
        except Exception as e:
            out[name]["adjoint_path"] = {"error": repr(e)}
        del h, tape, gh, y0, a
        torch.cuda.empty_cache()
    return out


def ensemble_extra(dev):
    """C4 (BASELINE configs[3]): M ensemble members / restarts with their OWN ml_net + expert parameters integrated and
    differentiated in ONE launch each way (odeint_ensemble); members shard over GPUs with no cross-member reduction."""
    import hybrid_ode_neurips_2021_b200 as H

    M, Bm, Dd = 64, 4096, 6
    torch.manual_seed(666)
    fs = [H.RocheODE(Dd, 1, T_MAX, 1, device=dev) for _ in range(M)]
    g = torch.Generator(device=dev).manual_seed(7)
    y0 = torch.empty(M * Bm, Dd, device=dev).exponential_(100.0, generator=g)
    a = torch.zeros(T, Bm, 1, device=dev)
    a[torch.randint(0, T_MAX, (Bm,), device=dev, generator=g), torch.arange(Bm, device=dev), 0] = \
        torch.rand(Bm, device=dev, generator=g) * 10 + 1e-3
    ens = H.EnsembleParams(fs)  # all members' parameters as one [M, P] leaf: one gradient tensor, no per-member packing
    tt = torch.arange(0, T_MAX + 1, 1, device=dev, dtype=torch.float32)
    W = torch.randn(T, M * Bm, Dd, device=dev, generator=g)

    def step():
        ens.flat.grad = None
        ens.set_action(a)
        z = y0.detach().requires_grad_(True)
        h = H.odeint_ensemble(ens, z, tt, method="rk4", options={"step_size": STEP, "expert_grads": False})
        (h * W).sum().backward()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_.record()
    for _ in range(3):
        step()
    e_.record(); torch.cuda.synchronize()
    ms = s_.elapsed_time(e_) / 3
    return {"value": M * Bm * N_STEPS / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "members": M, "patients_per_member": Bm,
            "latent_dim": Dd, "solver": "rk4(3/8) h=1/16", "note": "set_action + odeint_ensemble(EnsembleParams) fwd + reverse sweep "
            "through autograd, one parameter set per member (shared-memory staging instead of the constant bank), all "
            "members' gradients in one [M, P] tensor"}


def c5_extras(dev, fma_peak):
    """C5 (BASELINE configs[4]): the residual variant's vector field (run_simulation_residual.py trains the NeuralODE field)
    and the real-data decoder (run_real.py inputs: 48 hourly steps, 24 observed variables, 11 static covariates, latent 20)
    on a synthetic ICU-shaped cohort, forward + backward through the public API."""
    import hybrid_ode_neurips_2021_b200 as H

    def timed(fwd, n=5):
        """Median over n iterations of the forward and backward device times (a caching-allocator refill in one iteration
        must not masquerade as kernel time)."""
        for _ in range(3):
            fwd().backward()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf, tb = [], []
        for _ in range(n):
            torch.cuda.synchronize()
            ev[0].record(); loss = fwd(); ev[1].record(); loss.backward(); ev[2].record()
            torch.cuda.synchronize()
            tf.append(ev[0].elapsed_time(ev[1])); tb.append(ev[1].elapsed_time(ev[2]))
        return sorted(tf)[n // 2], sorted(tb)[n // 2]

    out = {}
    g = torch.Generator(device=dev).manual_seed(11)
    # --- NeuralODE field (Linear(D+1, 10D) -> Tanh -> Linear(10D, D) -> Tanh), rk4 on the headline grid ---------------------------
    Dn, Bn = 6, 131072
    torch.manual_seed(666)
    f = H.NeuralODE(Dn, 1, T_MAX, 1, device=dev)
    a = torch.zeros(T, Bn, 1, device=dev)
    a[torch.randint(0, T_MAX, (Bn,), device=dev, generator=g), torch.arange(Bn, device=dev), 0] = \
        torch.rand(Bn, device=dev, generator=g) * 10 + 1e-3
    f.set_action(a)
    y0 = (torch.rand(Bn, Dn, device=dev, generator=g) * 0.5).requires_grad_(True)
    tt = torch.arange(0, T_MAX + 1, 1, device=dev, dtype=torch.float32)
    W = torch.randn(T, Bn, Dn, device=dev, generator=g)

    def fwd_neural():
        f.zero_grad(set_to_none=True); y0.grad = None
        return (H.odeint(f, y0, tt, method="rk4", options={"step_size": STEP}) * W).sum()

    tf, tb = timed(fwd_neural)
    ff = 2 * (Dn + 1) * 10 * Dn + 10 * Dn + 2 * 10 * Dn * Dn + Dn  # two mat-vecs (FMA = 2) + one flop per tanh
    step_flops = 4 * ff + 18 * Dn
    out["C5_residual_neuralode_dim6"] = {
        "value": Bn * N_STEPS / ((tf + tb) * 1e-3), "unit": UNIT, "fwd_ms": tf, "bwd_ms": tb, "patients": Bn, "latent_dim": Dn,
        "solver": "rk4(3/8) h=1/16", "roofline": {
            "bound": "fp32_fma", "peak": fma_peak, "unit": "TFLOP/s", "flops_per_traj_step_fwd": step_flops,
            "fwd": {"achieved": Bn * N_STEPS * step_flops / (tf * 1e-3) / 1e12,
                    "frac": Bn * N_STEPS * step_flops / (tf * 1e-3) / 1e12 / fma_peak},
            "bwd": {"achieved": Bn * N_STEPS * 3 * step_flops / (tb * 1e-3) / 1e12,
                    "frac": Bn * N_STEPS * 3 * step_flops / (tb * 1e-3) / 1e12 / fma_peak}},
        "note": "odeint(NeuralODE) + autograd backward (reverse sweep with hidden-unit-owned weight gradients); times include "
                "the host-side op overhead"}
    del f, a, y0, W
    # --- DecoderReal (hybrid: RocheODEReal = learned expert nets + GRU-ODE latent block), midpoint, perturb=True -------------------
    Tr, obs, Hd, t0, Z, Br = 48, 24, 43, 24, 20, 65536
    torch.manual_seed(666)
    dec = H.DecoderReal(obs, Z, 1, 11, Hd, Tr, 1.0, t0=t0, method="midpoint", ode_step_size=1.0, ode_type="hybrid", device=dev)
    ar = (torch.rand(Tr, Br, 1, device=dev, generator=g) < 0.2).float() * torch.rand(Tr, Br, 1, device=dev, generator=g)
    sr = torch.randn(Tr, Br, 11, device=dev, generator=g)
    xr = torch.randn(Tr, Br, obs, device=dev, generator=g)
    mr = (torch.rand(Tr, Br, obs, device=dev, generator=g) < 0.5).float()
    z0 = (torch.randn(Br, Z, device=dev, generator=g) * 0.1).requires_grad_(True)

    def fwd_real():
        dec.zero_grad(set_to_none=True); z0.grad = None
        x_hat, _ = dec(z0, ar, sr)
        return torch.sum((xr[t0:] - x_hat) ** 2 * mr[t0:]) / Br  # model.py:1247

    tf, tb = timed(fwd_real)
    n_steps = dec.t.numel() - 1
    out["C5_icu_decoder_real_hybrid"] = {
        "value": Br * n_steps / ((tf + tb) * 1e-3), "unit": UNIT, "fwd_ms": tf, "bwd_ms": tb, "patients": Br,
        "patients_per_s": Br / ((tf + tb) * 1e-3), "latent_dim": Z, "hidden_dim": Hd, "obs_dim": obs, "t_max": Tr, "t0": t0,
        "solver": "midpoint h=1 (perturb=True), {} steps".format(n_steps),
        "note": "DecoderReal.forward (set_action_static + solve) + Linear-ELU-Linear read-out and masked SSE in ATen + backward"}
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
    ap.add_argument("--patients", type=int, default=1 << 20, help="patients per GPU (weak scaling) or in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
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
