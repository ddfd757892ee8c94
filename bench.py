#!/usr/bin/env python
"""bench.py -- inner-loop solves/sec of the ICNN argmin path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl reference]

One "step" = one solveBatch over one minibatch of synthetic input (default workload C2 =
BASELINE.json configs[1]: Olivetti-completion dims, n_y=2048, batch 400, 30 bundle iterations).
metric = B x iterations-executed / seconds ("a solve" = one sample advanced one inner iteration;
iterations-executed honours the reference's early return when every sample has finished,
lib/bundle_entropy.py:239).  Prints ONE JSON line (rank 0).

  value        device-resident: gates + y0 already in HBM, CUDA-event time of the fused loop
  e2e          same metric through icnn_b200.bundle_entropy.solveBatch with HOST buffers
               (H2D of x and y0, x-path gate precompute, loop, D2H of y*) inside the timed region
  roofline     dominant kernel class, CUDA-event time measured live in an instrumented pass
  cpu_baseline the numpy oracle port of the reference's solveBatch on the host cores (bounded sample)

--impl reference: times the CPU implementation only (oracle port; the reference itself is
Python and /root/reference does not exist on the GPU box), all host cores, same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "inner-loop solves/sec (batch x iters)"
UNIT = "solves/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            pk = json.load(f)
        return dict(hbm_gbs=float(pk["hbm_gbs"]), bf16_tflops=float(pk["bf16_tflops"]),
                    bf16_sustained=float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0,
                source="fallback (B200_PROFILING.md)")


def flop_fg(cfg):
    """Algorithmic FLOPs of one f/grad row: 4 * (n * sum s_i + sum s_{i-1} s_i)."""
    s = list(cfg["hidden"]) + [1]
    mac = cfg["n"] * sum(s) + sum(s[i - 1] * s[i] for i in range(1, len(s)))
    return 4.0 * mac


# ------------------------------------------------------------------------------------------
# CPU side (oracle port) -- the only place bench.py executes oracle/
# ------------------------------------------------------------------------------------------

def _cpu_worker(args):
    workload, seed, lo, hi, Bgen, nIter = args
    from threadpoolctl import threadpool_limits
    from oracle import bundle_np, picnn_np
    from icnn_b200 import workloads
    cfg = workloads.CONFIGS[workload]
    p, x, y0 = workloads.make_inputs(workload, B=Bgen, seed=seed)
    x, y0 = x[lo:hi], y0[lo:hi].copy()
    # float32 arithmetic mimics the reference's TF-backed fg; the values are handed over in float64
    # arrays so that the CPU arm does the same work as the GPU arm: np.linalg.matrix_rank scales its
    # tolerance with the row dtype, and with float32-typed rows most samples hit the rank stop early
    # (they would still be counted as "solves" while doing nothing)
    fg = picnn_np.make_fg(p, x, dtype=np.float32, out_dtype=np.float64, affine=cfg["affine"])
    iters = [0]

    def cb(t, *a):
        iters[0] = t + 1

    with threadpool_limits(limits=1), np.errstate(all="ignore"):
        t0 = time.perf_counter()
        bundle_np.solve_batch(fg, y0, nIter=nIter, variant=cfg["variant"], callback=cb)
        dt = time.perf_counter() - t0
    return (hi - lo) * iters[0], dt


def usable_cores():
    """Host cores this process may really use: min(affinity mask, cgroup v2/v1 CPU quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path, parse in (("/sys/fs/cgroup/cpu.max", lambda t: t.split()),
                        ("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", lambda t: [t.strip(), open(
                            "/sys/fs/cgroup/cpu/cpu.cfs_period_us").read().strip()])):
        try:
            q, per = parse(open(path).read())
            if q != "max" and int(q) > 0:
                n = min(n, max(1, int(int(q) / int(per))))
            break
        except Exception:
            continue
    return max(1, n)


_POOL = {}


def _close_pools():
    for pl in _POOL.values():
        pl.terminate()
        pl.join()
    _POOL.clear()


def _pool(procs):
    import multiprocessing as mp
    if procs not in _POOL:
        # spawn (not fork): the parent may hold an initialised CUDA context and torch threads
        _POOL[procs] = mp.get_context("spawn").Pool(procs)
    return _POOL[procs]


def cpu_reference(workload, rows, procs, seed=None, nIter=None):
    """Oracle-port solveBatch on ``rows`` rows of the workload, split over ``procs`` processes
    (samples are independent, lib/bundle_entropy.py:211).  Returns (solves, seconds) where
    seconds = the slowest worker's solveBatch time (process start-up / input generation are
    not counted)."""
    from icnn_b200 import workloads
    cfg = workloads.CONFIGS[workload]
    nIter = cfg["nIter"] if nIter is None else nIter
    seed = cfg["seed"] if seed is None else seed
    rows = max(1, min(rows, cfg["B"]))
    procs = max(1, min(procs, rows))
    bounds = [(rows * i) // procs for i in range(procs + 1)]
    jobs = [(workload, seed, bounds[i], bounds[i + 1], rows, nIter) for i in range(procs)]
    res = _pool(procs).map(_cpu_worker, jobs, chunksize=1)
    return sum(r[0] for r in res), max(r[1] for r in res)


def cpu_baseline(workload, budget_s=15.0):
    """Bounded sample sized for ~budget_s seconds on all host cores."""
    from icnn_b200 import workloads
    cfg = workloads.CONFIGS[workload]
    procs = usable_cores()
    # calibrate on a tiny sample (one row per process, few iterations are not representative:
    # cost grows with the bundle, so calibrate with the full iteration count on 1 row/proc)
    cal_rows = min(cfg["B"], procs)
    s0, w0 = cpu_reference(workload, cal_rows, procs)
    per_row = w0 / max(1, (cal_rows + procs - 1) // procs)
    rows = int(max(cal_rows, min(cfg["B"], procs * max(1.0, (budget_s / max(per_row, 1e-3))))))
    if rows > cal_rows:
        solves, wall = cpu_reference(workload, rows, procs)
    else:
        solves, wall, rows = s0, w0, cal_rows
    return dict(value=solves / wall, unit=UNIT, cores=min(procs, rows), kind="port",
                sample="%d of %d rows x %d iterations requested (%d solves executed, %.1f s wall), "
                       "numpy oracle port of lib/bundle_entropy.solveBatch(solver='pc'), float32-arithmetic fg, "
                       "%d processes x 1 BLAS thread" % (rows, cfg["B"], cfg["nIter"], solves, wall,
                                                         min(procs, rows)))


# ------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.reasons, self.max = index, [], set(), None
        self._stop = threading.Event()
        self._th = None

    def _run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max, reasons=sorted(self.reasons), samples=len(self.samples))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------

def run_gpu(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import icnn_b200
    from icnn_b200 import _capi, bundle_entropy, workloads
    from icnn_b200 import dist as idist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if rank == 0 and world == 1 and args.gpus > 1:
            print("bench.py: --gpus %d needs torchrun (WORLD_SIZE=1 seen)" % args.gpus, file=sys.stderr)
            sys.exit(2)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workloads.CONFIGS[args.workload]
    B, n, nIter = cfg["B"], cfg["n"], cfg["nIter"]
    strong = (args.scaling == "strong")
    if strong:
        # strong scaling: the workload's B rows are sharded over the ranks (contiguous blocks,
        # icnn_b200/dist.py) -- e.g. C4: 65 536 replay samples over 8 GPUs (BASELINE.json configs[3])
        p, x_all, y0_all = workloads.make_inputs(args.workload)
        lo, hi = idist.shard_rows(B, rank, world)
        x, y0 = x_all[lo:hi], y0_all[lo:hi]
        Btot, B = B, hi - lo
        p0 = p
    else:
        # weak scaling: every rank solves its own B rows (different seed -> different rows)
        p, x, y0 = workloads.make_inputs(args.workload, seed=cfg["seed"] + 7919 * rank)
        p0 = workloads.make_inputs(args.workload, B=1)[0] if rank else p   # theta replicated = rank 0's
    net = icnn_b200.PICNN.from_params(p0, device=dev)
    x_pin = torch.from_numpy(x.astype(np.float32)).pin_memory()
    y0_pin = torch.from_numpy(y0).pin_memory()
    variant, solver = cfg["variant"], args.solver
    KS = (nIter if variant == "rl" else min(nIter, n)) + 1
    ccfg = bundle_entropy._make_cfg(variant, solver, nIter, None, None, 0, n, KS)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    Bglob = Btot if strong else B * world
    y_all = torch.empty(Bglob, n, dtype=torch.float64, device=dev) if world > 1 else None

    def gather(y_local):
        if strong:
            y_all.copy_(idist.allgather_rows(y_local, Bglob))
        else:
            dist.all_gather_into_tensor(y_all, y_local)

    # ---- device-resident step -----------------------------------------------------------------
    fg = net.bind(x_pin.to(dev), affine=cfg["affine"])
    st = bundle_entropy.BundleState(B, n, KS, dev, keep_xs=True, nIter=nIter)
    y0_dev = y0_pin.to(dev)

    def step_device():
        st.y.copy_(y0_dev)
        _capi.check(_capi.lib.icnn_solve_batch_fused(net._h, C.byref(fg.c_gates), C.byref(ccfg), C.byref(st.c),
                                                     fg.ws.data_ptr(), stream))
        if world > 1:
            gather(st.y)

    def iters_executed():
        na = st.nactive.cpu().numpy()
        return int(np.sum(na[:nIter] > 0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()
    its = iters_executed()
    if world > 1:   # a job-level count: the slowest rank's shard decides when the loop is over
        t_its = torch.tensor([its], dtype=torch.int64, device=dev)
        dist.all_reduce(t_its, op=dist.ReduceOp.MAX)
        its = int(t_its.item())
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for s in range(args.steps):
            flush.fill_(s & 0xFF)           # L2 flush between timed iterations (untimed)
            ev[s][0].record()
            step_device()
            ev[s][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms = float(tmax.item())
    ms_per_step = dev_ms / args.steps
    value = Bglob * its / (ms_per_step * 1e-3)

    # ---- end to end through the public API, host buffers -------------------------------------
    def step_e2e():
        fg_h = net.bind(x_pin, affine=cfg["affine"])            # H2D x + x-path gate precompute
        out = bundle_entropy.solveBatch(fg_h, y0_pin, nIter=nIter, solver=solver,
                                        variant=variant, return_state=True)   # H2D y0 (pinned) ... D2H y*
        if world > 1:
            gather(out[-1].y)
        return out

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        out = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    its_e2e = int(np.sum(out[-1].nactive.cpu().numpy()[:nIter] > 0))
    e2e = dict(value=Bglob * its_e2e / e2e_s, unit=UNIT, ms_per_step=e2e_s * 1e3,
               h2d_bytes_per_step=int(x_pin.numel() * 4 + y0_pin.numel() * 8),
               d2h_bytes_per_step=int(B * n * 8 + B * 4 * 2 + (nIter + 1) * 4))

    # ---- instrumented pass: per-kernel-class CUDA-event time (K1 = PICNN f/grad, K2 = bundle step)
    reps = 3
    k1_ms = k2_ms = 0.0
    for _ in range(reps):
        st.y.copy_(y0_dev)
        _capi.check(_capi.lib.icnn_bundle_init(C.byref(st.c), nIter, stream))
        evs = []
        for t in range(its):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            _capi.check(_capi.lib.icnn_picnn_fg(net._h, C.byref(fg.c_gates), st.y32.data_ptr(), st.f.data_ptr(),
                                                st.G.data_ptr(), 0, st.perm.data_ptr(), st.count.data_ptr(), KS,
                                                fg.ws.data_ptr(), None, stream))
            e1.record()
            _capi.check(_capi.lib.icnn_bundle_step(C.byref(ccfg), C.byref(st.c), t, stream))
            e2.record()
            evs.append((e0, e1, e2))
        torch.cuda.synchronize()
        k1_ms += sum(a.elapsed_time(b) for a, b, _ in evs) / reps
        k2_ms += sum(b.elapsed_time(c) for _, b, c in evs) / reps
    ksum = int(st.ksum.sum().item())
    solves_local = B * its
    peaks = load_peaks()
    # K2 algorithmic bytes: per solve 4n(k_t + 2) + 8 k_t, summed exactly via ksum
    nsteps_k2 = int(st.newton_its.numel())  # noqa: F841
    k2_bytes = 4.0 * n * (ksum + 2.0 * solves_local) + 8.0 * ksum
    k1_flops = flop_fg(cfg) * solves_local
    roof_k2 = dict(kernel="bundle_step_kernel", bound="hbm", achieved=k2_bytes / (k2_ms * 1e-3) / 1e9,
                   peak=peaks["hbm_gbs"], unit="GB/s", traffic=None, launches=its,
                   ms_per_launch=k2_ms / max(its, 1), share_of_step=k2_ms / (k1_ms + k2_ms))
    roof_k2["frac"] = roof_k2["achieved"] / roof_k2["peak"]
    if args.workload == "C2":
        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
        # (profiles/r01_c2_summary.md section 3: outer iteration t = 20, k ~ 19 rows; the algorithmic
        # bytes of that launch are 69 MB)
        roof_k2["traffic"] = 72.18e6 + 9.43e6
        roof_k2["traffic_source"] = "ncu capture at t=20 (profiles/r01_c2_summary.md); per-launch, bytes"
    tc_shape = B >= 64    # every width runs on the tcgen05 path (operands are pitch-padded in the library)
    tc_on = tc_shape and not os.environ.get("ICNN_K1", "tc").startswith("s")
    k1_name = ("tc_gemm_kernel (tcgen05 3xTF32 + TMA) + gate_y + out_layer" if tc_on
               else "gated_gemm_kernel + out_layer (FP32 FFMA)")
    roof_k1 = dict(kernel=k1_name, bound="tensor",
                   achieved=k1_flops / (k1_ms * 1e-3) / 1e12, peak=peaks["bf16_sustained"], unit="TFLOP/s",
                   traffic=None, launches=its * (2 * len(cfg["hidden"]) + 2),
                   ms_per_launch=k1_ms / max(its * (2 * len(cfg["hidden"]) + 2), 1),
                   share_of_step=k1_ms / (k1_ms + k2_ms))
    roof_k1["frac"] = roof_k1["achieved"] / roof_k1["peak"]
    dominant = roof_k2 if k2_ms >= k1_ms else roof_k1
    dominant = dict(dominant, peak_source=peaks["source"])

    # ---- secondary: the north-star target shape T (batch 4096 / n_y 512), device-resident ----
    extra_T = None
    if args.workload != "T" and not args.no_target_shape:
        cT = workloads.CONFIGS["T"]
        pT, xT, y0T = workloads.make_inputs("T", seed=cT["seed"] + 7919 * rank)
        netT = icnn_b200.PICNN.from_params(pT, device=dev)
        fgT = netT.bind(torch.from_numpy(xT.astype(np.float32)).to(dev))
        KST = min(cT["nIter"], cT["n"]) + 1
        cfT = bundle_entropy._make_cfg("lib", solver, cT["nIter"], None, None, 0, cT["n"], KST)
        stT = bundle_entropy.BundleState(cT["B"], cT["n"], KST, dev, keep_xs=True, nIter=cT["nIter"])
        y0T_dev = torch.from_numpy(y0T).to(dev)

        def step_T():
            stT.y.copy_(y0T_dev)
            _capi.check(_capi.lib.icnn_solve_batch_fused(netT._h, C.byref(fgT.c_gates), C.byref(cfT), C.byref(stT.c),
                                                         fgT.ws.data_ptr(), stream))
        for _ in range(3):
            step_T()
        torch.cuda.synchronize()
        evT = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a_, b_ in evT:
            flush.fill_(1)
            a_.record(); step_T(); b_.record()
        torch.cuda.synchronize()
        msT = sum(a_.elapsed_time(b_) for a_, b_ in evT) / len(evT)
        itsT = int(np.sum(stT.nactive.cpu().numpy()[:cT["nIter"]] > 0))
        extra_T = {"workload": "T: m=512 n_y=512 hidden=[1024,1024] batch=4096 nIter=10 (north-star target shape)",
                   "value": cT["B"] * itsT / (msT * 1e-3), "unit": UNIT, "ms_per_step": msT, "iters_executed": itsT,
                   "n_gpus": 1}
        del stT, fgT, netT
    # ---- secondary (C3 only; SURVEY.md section 8d config 3 asks for both inner loops): the 30-step
    # momentum-GD loop (multi-label-cls/icnn-back.py:36-38 defaults) next to the bundle loop, the final
    # mean f(y) - H(y) of each (what ebundle-vs-gd.py:94-99 plots), and the GD training backward ----
    extra_gd = None
    if args.workload == "C3" and world == 1 and not cfg["affine"]:
        from icnn_b200 import gd as _gd, gd_grad as _gdg

        def f_minus_h(y32):
            f_, _ = fg.fg_device(y32.contiguous())
            ent = -(y32 * torch.log(y32) + (1 - y32) * torch.log(1 - y32))
            ent = torch.nan_to_num(ent, nan=0.0).sum(1)        # 0 log 0 = 0 (ebundle-vs-gd.py:38-41)
            return float((f_ - ent).mean())

        y0f = y0_dev.to(torch.float32)
        tYd = (torch.rand(B, n, device=dev, generator=torch.Generator(device=dev).manual_seed(5)) < 0.1).float()

        def timed(fn, reps=5):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0_.record()
            for _ in range(reps):
                r_ = fn()
            e1_.record(); torch.cuda.synchronize()
            return e0_.elapsed_time(e1_) / reps, r_

        ms_gd, (y_gd, _f) = timed(lambda: _gd.solve(fg, y0f, nIter=30, lr=0.01, momentum=0.3, return_device=True))
        ms_bw, _r = timed(lambda: _gdg.gd_grad(fg, y0f, tYd, nIter=30, lr=0.01, momentum=0.3, return_device=True))
        step_device(); torch.cuda.synchronize()
        extra_gd = {"gd_inner_loop": {"iters": 30, "lr": 0.01, "momentum": 0.3, "ms": ms_gd,
                                      "value": B * 30 / (ms_gd * 1e-3), "unit": UNIT,
                                      "mean_f_minus_H": f_minus_h(y_gd)},
                    "bundle_inner_loop": {"iters": its, "mean_f_minus_H": f_minus_h(st.y.to(torch.float32))},
                    "gd_training_backward_ms": ms_bw}
    if world > 1:
        dist.barrier()
    line = None
    if rank == 0:
        cpub = None
        if not args.no_cpu_baseline:
            cpub = cpu_baseline(args.workload, args.cpu_seconds)
        tc_path = k1_name.startswith("tc_gemm")
        launches_per_step = 2 + nIter * (2 * len(cfg["hidden"]) + (3 if tc_path else 2))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 (K1 PICNN f/grad) + f64 (K2 bundle solve)",
            "data": "synthetic", "impl": "icnn_b200",
            "config": {"workload": "%s: m=%d n_y=%d hidden=%s batch=%d/GPU nIter=%d variant=%s solver=%s"
                                   % (args.workload, cfg["m"], n, cfg["hidden"], B, nIter, variant, solver),
                       "global_batch": Bglob, "iters_executed": its, "iters_requested": nIter,
                       "parallelism": "sample-sharded x%d, one all-gather of y*" % world,
                       "l2": "512 MiB buffer written between timed steps (L2 flush)",
                       "wall_s_timed_region": t_wall},
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "clocks": clk.summary(), "roofline": dominant,
            "kernels": {"K1_picnn_fg": roof_k1, "K2_bundle_step": roof_k2},
            "cpu_baseline": cpub, "target_shape": extra_T, "gd_mode": extra_gd,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def run_reference(args):
    """The reference's CPU path (oracle port, all host cores) on the same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from icnn_b200 import workloads
    cfg = workloads.CONFIGS[args.workload]
    procs = usable_cores()
    # bounded sample per step, sized from one calibration run so K+W steps end within minutes
    cal_rows = min(cfg["B"], procs)
    s0, w0 = cpu_reference(args.workload, cal_rows, procs)
    total_steps = args.steps + args.warmup
    per_step_budget = max(2.0, min(20.0, 150.0 / max(total_steps, 1)))
    per_row = w0 / max(1, (cal_rows + procs - 1) // procs)
    rows = int(max(cal_rows, min(cfg["B"], procs * max(1.0, per_step_budget / max(per_row, 1e-3)))))
    for _ in range(args.warmup):
        cpu_reference(args.workload, rows, procs)
    tot_s, tot_w = 0, 0.0
    for _ in range(args.steps):
        s, w = cpu_reference(args.workload, rows, procs)
        tot_s += s
        tot_w += w
    value = tot_s / tot_w
    sample = ("%d of %d rows x nIter=%d per step, numpy oracle port of lib/bundle_entropy.solveBatch"
              "(solver='pc') with float32-arithmetic fg, %d processes x 1 BLAS thread" % (rows, cfg["B"], cfg["nIter"],
                                                                              min(procs, rows)))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot_w / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 solver / f32 fg", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": "%s: m=%d n_y=%d hidden=%s batch=%d nIter=%d variant=%s solver=pc"
                                   % (args.workload, cfg["m"], cfg["n"], cfg["hidden"], cfg["B"], cfg["nIter"],
                                      cfg["variant"])},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(procs, rows), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="icnn_b200", choices=["icnn_b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5", "T"])
    ap.add_argument("--solver", default="pc", choices=["pc", "newton"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's batch per GPU (default); strong: the batch sharded over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-target-shape", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_gpu(args)
    finally:
        _close_pools()


if __name__ == "__main__":
    main()
