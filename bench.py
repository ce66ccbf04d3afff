#!/usr/bin/env python
"""bench.py -- FastPM forward+vjp steps/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--nmesh 128] [--nsteps 10]
    python bench.py --impl reference ...        # the CPU path (numpy/C oracle) on the host cores

Workload (BASELINE.json configs[1]): white noise -> induce_correlation -> ``nbody`` with
``nsteps`` FastPM steps (``len(stages) = nsteps + 1``) on ``nmesh^3`` particles / mesh, fp64;
one bench "step" = ONE ``Model.compute_with_vjp`` of that graph (forward + reverse sweep with a
cotangent on ``dx``).  ``value`` = nsteps / (seconds per bench step), inputs resident in HBM;
``e2e`` = the same call with HOST numpy inputs (white noise + cotangent copied host->device
inside the timed region, the gradient copied back device->host).

Timing: W untimed warm-up steps, then K steps each bracketed by CUDA events on the stream the
kernels are launched on, an L2 flush (256 MiB write) between steps outside the events, a
barrier + synchronize on both sides, max over ranks.

N > 1 (torchrun, one rank per GPU): ONE problem, slab-decomposed over the N GPUs (mesh planes
and the particles on them per rank; particle exchange and FFT transposes are stores into the
peers' memory over NVLink -- vmad_b200/csrc/peer.cu -- with NCCL all-to-all as the fallback).  Weak scaling: the global mesh grows with N so that every GPU
keeps nmesh^3 cells and particles ([nmesh*a, nmesh*b, nmesh*c] with a*b*c = N, cell size
fixed); ``value`` = N * nsteps / seconds, i.e. steps/s in units of the N = 1 problem.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METRIC = "FastPM forward+vjp steps/s"
UNIT = "steps/s"


def workload_name(nmesh, nsteps):
    return ("FastPM %d-step forward+vjp, %d^3 particles / %d^3 mesh, fp64, white noise -> "
            "induce_correlation -> nbody (BASELINE.json configs[1] shape)" % (nsteps, nmesh, nmesh))


class Cosmology(object):
    Om0 = 0.3075


def power(k):
    k = numpy.where(k == 0, 1.0, k)
    return 1.5e6 * k / (1.0 + (k / 0.05) ** 2) ** 2


def global_shape(nmesh, world):
    """weak scaling: distribute the factors of ``world`` (a power of two) over the axes"""
    dims = [nmesh, nmesh, nmesh]
    w, d = world, 0
    while w > 1:
        if w % 2:
            raise SystemExit("bench.py: --gpus must be a power of two")
        dims[d % 3] *= 2
        w //= 2
        d += 1
    return dims


def make_inputs(nmesh, seed=None, dims=None):
    """synthetic Gaussian initial conditions (host arrays, global)"""
    dims = [nmesh] * 3 if dims is None else list(dims)
    ntot = dims[0] * dims[1] * dims[2]
    rng = numpy.random.default_rng(300 + nmesh if seed is None else seed)
    wn = numpy.fft.rfftn(rng.standard_normal(tuple(dims))) / ntot ** 0.5
    cot = rng.standard_normal((ntot, 3))
    return wn, cot


def build_model(pm, q, stages):
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    with Builder() as m:
        x = m.input('x')
        rhok = fastpm.induce_correlation(x, power, pm)
        dx, p, f = fastpm.nbody(rhok, q, stages, Cosmology, pm)
        m.output(dx=dx, p=p, f=f)
    return m


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock + clock-event reasons sampled DURING the timed region.

    NVML is read in-process from a background thread (a few cheap driver queries every 20 ms);
    an external ``nvidia-smi -lms`` loop is only the fallback -- its heavier polling perturbs
    multi-process runs whose ranks spin on each other's flags.
    """
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.sm, self.mx, self.reasons = [], [], set()

    def _nvml_loop(self, nv, handle):
        masks = [('hw_slowdown', nv.nvmlClocksEventReasonHwSlowdown),
                 ('hw_thermal_slowdown', nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ('sw_thermal_slowdown', nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ('sw_power_cap', nv.nvmlClocksEventReasonSwPowerCap)]
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))
                for name, m in masks:
                    if bits & m:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        mode = os.environ.get('VMAD_BENCH_CLOCKS', 'nvml')
        if mode == 'off':
            return
        try:
            if mode == 'smi':
                raise RuntimeError
            import threading
            import pynvml as nv
            nv.nvmlInit()
            try:    # CUDA's device numbering may differ from NVML's (CUDA_VISIBLE_DEVICES)
                import torch
                uuid = "GPU-%s" % torch.cuda.get_device_properties(self.gpu_index).uuid
                handle = nv.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                handle = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)))
            self._stop = threading.Event()
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '500'],
                stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            if self.sm:
                out.update(sm_mhz=float(numpy.median(self.sm)), sm_max_mhz=float(max(self.mx)),
                           reasons=sorted(self.reasons), samples=len(self.sm), source='nvml')
            return out
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [l.strip().split(', ') for l in open(self.path) if l.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.strip().lower() == 'active':
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(numpy.median(sm)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), samples=len(sm), source='nvidia-smi')
        return out


# ---------------------------------------------------------------------------------------------
# CPU path (oracle): cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------------
def cpu_fwd_vjp_seconds(nmesh, nsteps, threads, repeats=1):
    """seconds per forward+vjp of the same graph on the numpy/C oracle backend"""
    import oracle
    oracle.activate(shims=False)
    from pmesh import pm as opm
    from pmesh.pm import ParticleMesh as OraclePM
    opm.set_cic_threads(threads)
    opm.set_fft_workers(threads)
    N = nmesh
    pm = OraclePM([N] * 3, 100.0 * N / 32)
    wn, cot = make_inputs(N)
    q = pm.generate_uniform_particle_grid(shift=0.0)
    stages = numpy.linspace(0.1, 1.0, nsteps + 1)
    m = build_model(pm, q, stages)
    x = pm.create('complex', value=wn)
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        m.compute_with_vjp(init=dict(x=x), v=dict(_dx=cot))
        ts.append(time.perf_counter() - t0)
    return ts


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    threads = host_threads()
    nmesh = min(args.nmesh, args.cpu_nmesh)
    ts = cpu_fwd_vjp_seconds(nmesh, args.nsteps, threads, repeats=args.warmup + args.steps)
    timed = ts[args.warmup:]
    sec = float(numpy.mean(timed))
    value = args.nsteps / sec
    sample = ("full %d-step forward+vjp at %d^3 on the numpy + C(OpenMP) oracle, %d threads"
              % (args.nsteps, nmesh, threads))
    if nmesh != args.nmesh:
        sample += " (bounded sample: the GPU arm runs %d^3)" % args.nmesh
    if args.gpus > 1:
        sample += ("; bounded sample of the %d-GPU weak-scaling workload: one %d^3 block of the global mesh "
                   "%s, throughput counted in the same %d^3-problem units as the GPU arm's value"
                   % (args.gpus, args.nmesh, global_shape(args.nmesh, args.gpus), args.nmesh))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.nmesh, args.nsteps), "nmesh": args.nmesh,
                   "nsteps": args.nsteps, "cpu_nmesh": nmesh},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        return float(json.load(open(path))['hbm_gbs']), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel):
    """DRAM bytes per launch of the roofline kernel from the committed ncu --set full capture"""
    try:
        d = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))
        return d.get(kernel)
    except Exception:
        return None


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from vmad_b200 import _capi as C
    from vmad_b200.pm import DeviceArray, ParticleMesh

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N = args.nmesh
    dims = global_shape(N, world)
    comm = None
    if world > 1:
        from vmad_b200.dist import TorchComm
        comm = TorchComm()
    pm = ParticleMesh(dims, [100.0 * n / 32 for n in dims], comm=comm)
    pm.force_recompute = N >= 512  # keep potk instead of 3 force meshes per step on the tape
    q = pm.generate_uniform_particle_grid(shift=0.0)
    npl = len(q)
    if int(numpy.prod(dims)) <= 512 ** 3:
        wn_g, cot_g = make_inputs(N, dims=dims)
        # this rank's share: complex slab along axis 1, particles of its planes (lattice order)
        wn_h = numpy.ascontiguousarray(wn_g[:, pm.cstart1:pm.cstart1 + pm.cshape[1]]) if world > 1 else wn_g
        cot_h = numpy.ascontiguousarray(cot_g[rank * npl:(rank + 1) * npl])
        del wn_g, cot_g
    else:
        # global arrays of this size do not belong on every rank's host: draw the local complex
        # slab and the local cotangent rows directly (unit-variance modes, per-rank streams)
        rng = numpy.random.default_rng(300 + N + 1000 * rank)
        wn_h = (rng.standard_normal(tuple(pm.cshape)) + 1j * rng.standard_normal(tuple(pm.cshape))) / 2 ** 0.5
        cot_h = rng.standard_normal((npl, 3))
    stages = numpy.linspace(0.1, 1.0, args.nsteps + 1)
    model = build_model(pm, q, stages)

    # resident inputs
    wn_d = pm.create('complex', value=wn_h)
    cot_d = DeviceArray.from_host(cot_h)
    # pinned host staging for the end-to-end leg
    wn_pin = torch.from_numpy(wn_h).pin_memory()
    cot_pin = torch.from_numpy(cot_h).pin_memory()
    grad_pin = torch.empty(wn_pin.shape, dtype=wn_pin.dtype).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def step_resident():
        (dx,), (g,) = model.compute_with_vjp(init=dict(x=wn_d), v=dict(_dx=cot_d))
        return dx, g

    def step_e2e():
        x = pm.create('complex')
        x.t.copy_(wn_pin, non_blocking=True)
        c = DeviceArray(cot_pin.to('cuda', non_blocking=True))
        (dx,), (g,) = model.compute_with_vjp(init=dict(x=x), v=dict(_dx=c))
        grad_pin.copy_(g.t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return dx, g

    # One GPU: nothing in the call waits for the device, so the whole forward+vjp (VM, ctypes,
    # ~700 launches) is captured once into a CUDA graph and each step is one graph launch on
    # the same input buffers (vmad_b200/cudagraph.py).  Across ranks the host needs the particle
    # counts of every decompose, so the call is not capturable and runs eagerly.
    use_graph = world == 1 and N < 512 and not args.no_graph
    eager_resident = step_resident
    launches_per_graph = 0
    if use_graph:
        from vmad_b200.cudagraph import GraphedCall
        for _ in range(2):
            eager_resident()
        l0 = C.launch_count()
        graphed = GraphedCall(eager_resident, warmup=1)
        launches_per_graph = (C.launch_count() - l0) // 2      # one warm-up call + the captured one

        def step_resident():                                   # noqa: F811
            return graphed()

        # end-to-end: the host<->device copies are nodes of a second captured graph; the upload
        # of the cotangent (50 MB) runs on a forked stream beside the forward sweep and joins
        # before the reverse sweep, which is the first reader (same calls as compute_with_vjp:
        # compute(return_tape=True) then tape.get_vjp().compute)
        side = torch.cuda.Stream()

        def e2e_call():
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                cot_d.t.copy_(cot_pin, non_blocking=True)
            wn_d.t.copy_(wn_pin, non_blocking=True)
            (dx,), tape = model.compute(['dx'], init=dict(x=wn_d), return_tape=True)
            cur.wait_stream(side)
            (g,) = tape.get_vjp().compute(tape.get_vjp_vout(), init=dict(_dx=cot_d))
            grad_pin.copy_(g.t, non_blocking=True)
            return dx, g

        graphed_e2e = GraphedCall(e2e_call, warmup=1)

        def step_e2e():                                        # noqa: F811
            dx, g = graphed_e2e()
            torch.cuda.current_stream().synchronize()
            return dx, g

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        # the graph, plans and layout caches built during warm-up are long-lived: move them out
        # of the cyclic collector's generations so that a full collection (tens of ms of host
        # time, which lock-stepped ranks all wait for) does not land inside the timed steps
        gc.collect()
        gc.freeze()
        barrier()
        l0 = C.launch_count()
        total_ms = 0.0
        per_step = []
        mallocs = [torch.cuda.memory_stats().get('num_device_alloc', 0)]
        wall0 = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            per_step.append(e0.elapsed_time(e1))
            total_ms += per_step[-1]
            mallocs.append(torch.cuda.memory_stats().get('num_device_alloc', 0))
        if rank == 0:
            print("[bench] %s per-step ms: %s; cudaMalloc calls per step: %s" % (
                fn.__name__, ["%.2f" % v for v in per_step],
                [b - a for a, b in zip(mallocs[:-1], mallocs[1:])]), file=sys.stderr)
        barrier()
        wall = time.perf_counter() - wall0
        launches = C.launch_count() - l0
        if use_graph and fn is not eager_resident:
            launches = launches_per_graph * steps       # replayed launches are not re-counted
        t = torch.tensor([total_ms], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, wall

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, launches, wall = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))

    e2e_check = None
    if use_graph:       # what came back over PCIe from the replayed pipeline == the eager gradient
        _, g_e = eager_resident()
        torch.cuda.synchronize()
        ref = g_e.t.cpu()
        e2e_check = float((grad_pin - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt())
        del ref, g_e
    eager_ms = None
    if use_graph:       # the same step without the graph, for the record
        eager_total, _, _ = timed(eager_resident, args.steps, 1)
        eager_ms = eager_total / args.steps
    ms_per_step = total_ms / args.steps
    value = world * args.nsteps / (ms_per_step * 1e-3)
    e2e_value = world * args.nsteps / (e2e_ms / args.steps * 1e-3)

    # ---- roofline of the dominant hand-written kernel, measured live with CUDA events on the
    # ---- launching stream, on the particle distribution the run ended with
    dx, _ = step_resident()
    x_final = q + dx
    lay = pm.decompose(x_final)
    xs = lay.exchange(x_final)
    Np, Nc = len(xs), int(numpy.prod(pm.rshape))
    peak, peak_src = measured_peak()

    def kernel_time(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return float(numpy.mean(ts))

    # (a) every paint launch of ONE more (untimed, otherwise identical) step, timed in place
    from vmad_b200 import pm as pm_mod
    pm_mod.PAINT_EVENTS = []
    eager_resident()
    torch.cuda.synchronize()
    every = [(e0.elapsed_time(e1) * 1e-3, nb, kn) for e0, e1, nb, kn in pm_mod.PAINT_EVENTS]
    pm_mod.PAINT_EVENTS = None
    # the roofline entry is the single-mesh scatter kernel (the one the committed ncu capture and
    # `traffic` describe); the fused three-mesh variant of the reverse sweep is listed beside it
    live = [(t, nb) for t, nb, kn in every if kn == 'k_paint_agg']
    live_s = sum(t for t, _ in live)
    live_bytes = sum(nb for _, nb in live)
    fused3 = [(t, nb) for t, nb, kn in every if kn != 'k_paint_agg']
    all_paint_s = sum(t for t, _, _ in every)
    # (b) isolated loops with an L2 flush between launches
    field = pm.paint(xs, 1.0)
    kernels = {}
    for name, fn, nbytes in [
            ('k_paint_agg', lambda: pm.paint(xs, 1.0), 24.0 * Np + 8.0 * Nc),
            ('k_readout', lambda: field.readout(xs), 32.0 * Np + 8.0 * Nc),
            ('k_readout_grad', lambda: pm.paint_vjp(field, xs, mass=1.0), 56.0 * Np + 8.0 * Nc),
            ('layout_build', lambda: pm._decompose_local(xs), 48.0 * Np),
            ('r2c', lambda: field.r2c(), 16.06 * Nc)]:
        t = kernel_time(fn)
        kernels[name] = dict(ms=t * 1e3, achieved_gbs=nbytes / t / 1e9, frac=nbytes / t / 1e9 / peak,
                             gparticles_s=Np / t / 1e9)
    dom = 'k_paint_agg'
    nl = max(len(live), 1)
    roofline = {"bound": "hbm", "kernel": dom,
                "achieved": live_bytes / live_s / 1e9 if live_s > 0 else None,
                "peak": peak, "unit": "GB/s",
                "frac": live_bytes / live_s / 1e9 / peak if live_s > 0 else None,
                "traffic": recorded_traffic(dom), "peak_source": peak_src,
                "how": ("all %d single-mesh paint launches (memset + k_paint_agg each) of one forward+vjp "
                        "step, CUDA events on the launching stream, in place" % len(live)),
                "launches": len(live), "avg_launch_ms": live_s / nl * 1e3,
                "algorithmic_bytes_per_launch": live_bytes / nl,
                "share_of_step": live_s * 1e3 / (eager_ms or ms_per_step),
                "k_paint_agg3": ({"launches": len(fused3),
                                  "avg_launch_ms": sum(t for t, _ in fused3) / len(fused3) * 1e3,
                                  "algorithmic_bytes_per_launch": sum(nb for _, nb in fused3) / len(fused3),
                                  "achieved": sum(nb for _, nb in fused3) / sum(t for t, _ in fused3) / 1e9,
                                  "frac": sum(nb for _, nb in fused3) / sum(t for t, _ in fused3) / 1e9 / peak,
                                  "what": "three meshes per launch (3 memsets + kernel), 48 B/particle + "
                                          "24 B/cell: positions and weights shared"}
                                 if fused3 else None),
                "all_paint_share_of_step": all_paint_s * 1e3 / (eager_ms or ms_per_step),
                "isolated_l2_flushed": kernels}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        nm = min(N, args.cpu_nmesh)
        sec = cpu_fwd_vjp_seconds(nm, args.nsteps, threads, repeats=1)[0]
        sample = ("one full %d-step forward+vjp at %d^3 on the numpy + C(OpenMP) oracle, "
                  "%d threads (%.1f s)" % (args.nsteps, nm, threads, sec))
        if nm != N:
            sample += "; bounded sample: the GPU arm runs %d^3" % N
        cpu_baseline = {"value": args.nsteps / sec, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": sample}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(N, args.nsteps), "nmesh": N, "nsteps": args.nsteps,
                   "global_mesh": dims, "np": int(numpy.prod(dims)),
                   "boxsize": [100.0 * n / 32 for n in dims],
                   "parallelism": ("slab decomposition over %d GPUs (mesh axis 0; %s "
                                   "for particle exchange and FFT transposes), %d^3 cells per GPU"
                                   % (world, "peer-memory stores over NVLink (CUDA IPC)"
                                      if pm._peer() is not None else "NCCL all-to-all", N))
                   if world > 1 else "1 gpu",
                   "l2": "256 MiB flush write between timed steps",
                   "step": "one Model.compute_with_vjp = %d FastPM steps forward + reverse" % args.nsteps,
                   "launch": ("CUDA graph replay of the captured compute_with_vjp (same input buffers; "
                              "%d kernel launches per replay)" % launches_per_graph) if use_graph
                   else "eager (python VM + ctypes, one launch per kernel)",
                   "eager_ms_per_step": eager_ms, "e2e_gradient_vs_eager_rel_l2": e2e_check},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(wn_h.nbytes + cot_h.nbytes),
                "d2h_bytes_per_step": int(wn_h.nbytes)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "host_wall_s_timed_region": wall,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry the one JSON line and nothing else: point fd 1 at stderr for the rest
    of the process (NCCL, cuFFT and friends print their banners from C), keep the real one"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--nmesh', type=int, default=128)
    ap.add_argument('--nsteps', type=int, default=10)
    ap.add_argument('--cpu-nmesh', type=int, default=128,
                    help='largest mesh the CPU arm / cpu_baseline leg will run')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='run every step eagerly through the VM')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
