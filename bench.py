#!/usr/bin/env python
"""bench.py -- FastPM forward+vjp steps/s at 512^3 on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--nmesh 512] [--nsteps 10]
    python bench.py --impl reference ...        # the CPU path (numpy/C oracle) on the host cores
    python bench.py --micro ...                 # CIC paint / readout microbenchmark (configs[2])

Workload (the metric's own configuration): white noise -> induce_correlation -> ``nbody`` with
``nsteps`` FastPM steps (``len(stages) = nsteps + 1``) on ``nmesh^3`` particles and an
``nmesh^3`` mesh, fp64; one bench "step" = ONE ``Model.compute_with_vjp`` of that graph
(forward + reverse sweep with a cotangent on ``dx``).  ``value`` = nsteps / (seconds per bench
step), inputs resident in HBM; ``e2e`` = the same call with HOST buffers: white noise and
cotangent copied host->device and BOTH results (displacement ``dx`` and gradient) copied
device->host inside the timed region.

N > 1 (torchrun, one rank per GPU): the SAME global nmesh^3 problem, slab-decomposed over the N
GPUs -- strong scaling, ``value`` = nsteps / seconds.  Particle exchange and FFT transposes are
stores into the peers' memory over NVLink (vmad_b200/csrc/peer.cu; NCCL all-to-all is the
fallback).  The same JSON line carries side records: ``side.weak_128`` (128^3 cells and
particles PER GPU, the round-1 configuration: BASELINE.json configs[1] at N = 1), ``parity``
(GPU vs the numpy/C oracle at 128^3 x 10 steps on one GPU; N-rank vs 1-rank results of the
main problem and of small chi-square / FastPMSimulation problems across ranks) and, on 8 GPUs,
``side.gn_step_512`` (configs[3]) and ``side.fastpm_1024`` (configs[4]).

Timing: W untimed warm-up steps, then K steps each bracketed by CUDA events on the stream the
kernels are launched on, an L2 flush (256 MiB write) between steps outside the events, a
barrier + synchronize on both sides, max over ranks.
"""
import argparse
import collections
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

METRIC_FMT = "FastPM forward+vjp steps/s at %d^3"
UNIT = "steps/s"
SEED = 300


def workload_name(nmesh, nsteps):
    return ("FastPM %d-step forward+vjp, %d^3 particles / %d^3 mesh, fp64, white noise -> "
            "induce_correlation -> nbody" % (nsteps, nmesh, nmesh))


def workload_config(nmesh, nsteps):
    """the workload description: identical in the GPU arm and the CPU (reference) arm"""
    return {"workload": workload_name(nmesh, nsteps), "nmesh": nmesh, "nsteps": nsteps,
            "global_mesh": [nmesh] * 3, "np": nmesh ** 3, "boxsize": [100.0 * nmesh / 32] * 3,
            "stages": "linspace(0.1, 1.0, nsteps + 1)", "Om0": Cosmology.Om0,
            "inputs": "seeded per-plane numpy white noise (default_rng([%d, plane])), cotangent on dx" % SEED,
            "l2": "GPU arm: a 256 MiB buffer is written between timed steps (L2 flush); every step also streams "
                  ">> 126 MB (the inputs alone are larger than L2)"}


class Cosmology(object):
    Om0 = 0.3075


def power(k):
    k = numpy.where(k == 0, 1.0, k)
    return 1.5e6 * k / (1.0 + (k / 0.05) ** 2) ** 2


def weak_shape(nmesh, world):
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


# ---------------------------------------------------------------------------------------------
# synthetic inputs: per-plane seeded streams, so that any rank layout draws the same global field
# ---------------------------------------------------------------------------------------------
def noise_planes(dims, plane0, nplanes, seed=SEED):
    """real white noise of mesh planes [plane0, plane0 + nplanes) of a global [dims] field"""
    out = numpy.empty((nplanes, dims[1], dims[2]))
    for i in range(nplanes):
        out[i] = numpy.random.default_rng([seed, plane0 + i]).standard_normal((dims[1], dims[2]))
    return out


def cotangent_rows(dims, plane0, nplanes, seed=SEED):
    """cotangent of dx for the lattice particles of planes [plane0, plane0 + nplanes)"""
    per = dims[1] * dims[2]
    out = numpy.empty((nplanes * per, 3))
    for i in range(nplanes):
        out[i * per:(i + 1) * per] = numpy.random.default_rng([seed + 7919, plane0 + i]).standard_normal((per, 3))
    return out


def make_inputs(nmesh, seed=SEED, dims=None):
    """host arrays (global): white noise in k-space (rfftn of the real noise / sqrt(Ntot)) and
    the cotangent -- what the oracle leg consumes; the device legs build the same field from the
    same planes with their own r2c"""
    dims = [nmesh] * 3 if dims is None else list(dims)
    ntot = dims[0] * dims[1] * dims[2]
    wn = numpy.fft.rfftn(noise_planes(dims, 0, dims[0], seed)) / ntot ** 0.5
    return wn, cotangent_rows(dims, 0, dims[0], seed)


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
def cpu_fwd_vjp(nmesh, nsteps, threads, repeats=1, keep_result=False):
    """the same graph on the numpy/C oracle backend: seconds per forward+vjp (and the result)"""
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
    ts, res = [], None
    for _ in range(repeats):
        t0 = time.perf_counter()
        (dx,), (g,) = m.compute_with_vjp(init=dict(x=x), v=dict(_dx=cot))
        ts.append(time.perf_counter() - t0)
        if keep_result:
            res = (numpy.asarray(dx), numpy.asarray(g))
    return ts, res


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample_text(nsteps, cpu_nmesh, nmesh, threads, sec=None):
    s = ("one full %d-step forward+vjp at %d^3 on the numpy + C(OpenMP) oracle, %d threads%s"
         % (nsteps, cpu_nmesh, threads, " (%.1f s)" % sec if sec is not None else ""))
    if cpu_nmesh != nmesh:
        s += ("; bounded sample of the %d^3 workload: %d^3 particles and cells (same cell size, same graph) = "
              "1/%d of the work, value = measured %d^3 steps/s x (%d/%d)^3, i.e. scaled linearly in "
              "particles and cells to the metric's unit (favourable to the CPU: its FFT grows as N log N "
              "and its caches stop helping)"
              % (nmesh, cpu_nmesh, (nmesh // cpu_nmesh) ** 3, cpu_nmesh, cpu_nmesh, nmesh))
    return s


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    threads = host_threads()
    nm = min(args.nmesh, args.cpu_nmesh)
    ts, _ = cpu_fwd_vjp(nm, args.nsteps, threads, repeats=args.warmup + args.steps)
    sec = float(numpy.mean(ts[args.warmup:]))
    scale = (float(nm) / args.nmesh) ** 3
    value = args.nsteps / sec * scale
    line = {
        "impl": "reference", "metric": METRIC_FMT % args.nmesh, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3 / scale, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.nmesh, args.nsteps),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": cpu_sample_text(args.nsteps, nm, args.nmesh, threads),
                         "measured_ms_per_sample": sec * 1e3, "sample_nmesh": nm},
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


def recorded_traffic(nmesh, kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of ``kernel`` from the
    committed ``ncu --set full`` capture of this bench at this mesh size, or None"""
    try:
        d = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))
        return d.get(str(nmesh), {}).get(kernel)
    except Exception:
        return None


class Env(object):
    """process-wide state of the GPU arm"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
        self.comm = None
        if self.world > 1:
            from vmad_b200.dist import TorchComm
            self.comm = TorchComm()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device='cuda')
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        t = self.torch.tensor(numpy.atleast_1d(numpy.asarray(v, dtype='f8')), dtype=self.torch.float64,
                              device='cuda')
        if self.world > 1:
            self.dist.all_reduce(t)
        return t.cpu().numpy()

    def log(self, *a):
        if self.rank == 0:
            print("[bench]", *a, file=sys.stderr, flush=True)


class Problem(object):
    """one FastPM forward+vjp problem on the communicator ``comm`` (None: this GPU alone)"""

    def __init__(self, env, dims, nsteps, comm='env', force_recompute=None, pinned=True, seed=SEED):
        import torch
        from vmad_b200.pm import DeviceArray, ParticleMesh, RealField
        self.env = env
        comm = env.comm if comm == 'env' else comm
        self.dims = list(dims)
        self.nsteps = nsteps
        pm = ParticleMesh(self.dims, [100.0 * n / 32 for n in self.dims], comm=comm)
        ncell_local = int(numpy.prod(pm.rshape))
        pm.force_recompute = (ncell_local >= 512 ** 3) if force_recompute is None else force_recompute
        self.pm = pm
        self.q = pm.generate_uniform_particle_grid(shift=0.0)
        self.npl = len(self.q)
        n0l = pm.rshape[0]
        # the rank's planes of the real noise -> device r2c -> its slab of the k-space white noise
        noise = RealField(pm, torch.from_numpy(noise_planes(self.dims, pm.rstart0, n0l, seed)).cuda())
        self.wn = pm._r2c(noise, pm.Ntot ** -0.5)
        del noise
        self.cot = DeviceArray(torch.from_numpy(cotangent_rows(self.dims, pm.rstart0, n0l, seed)).cuda())
        self.model = build_model(pm, self.q, numpy.linspace(0.1, 1.0, nsteps + 1))
        self.pinned = None
        if pinned:
            wn_pin = self.wn.t.cpu().pin_memory()
            cot_pin = self.cot.t.cpu().pin_memory()
            self.pinned = dict(wn=wn_pin, cot=cot_pin,
                               grad=torch.empty_like(wn_pin).pin_memory(),
                               dx=torch.empty_like(cot_pin).pin_memory())
            self.side_up = torch.cuda.Stream()
            self.side_down = torch.cuda.Stream()
        self.h2d_bytes = int(self.wn.t.numel() * 16 + self.cot.t.numel() * 8)
        self.d2h_bytes = int(self.wn.t.numel() * 16 + self.cot.t.numel() * 8)

    def step_resident(self):
        (dx,), (g,) = self.model.compute_with_vjp(init=dict(x=self.wn), v=dict(_dx=self.cot))
        return dx, g

    def e2e_call(self):
        """host buffers in, host buffers out.  The cotangent upload runs on a side stream beside
        the forward sweep (the reverse sweep is its first reader) and the displacement goes back
        on another side stream beside the reverse sweep; the same calls as compute_with_vjp
        (compute(return_tape=True), then tape.get_vjp().compute)."""
        torch = self.env.torch
        pin = self.pinned
        cur = torch.cuda.current_stream()
        self.side_up.wait_stream(cur)
        with torch.cuda.stream(self.side_up):
            self.cot.t.copy_(pin['cot'], non_blocking=True)
        self.wn.t.copy_(pin['wn'], non_blocking=True)
        (dx,), tape = self.model.compute(['dx'], init=dict(x=self.wn), return_tape=True)
        self.side_down.wait_stream(cur)
        with torch.cuda.stream(self.side_down):
            pin['dx'].copy_(dx.t, non_blocking=True)
        cur.wait_stream(self.side_up)
        (g,) = tape.get_vjp().compute(tape.get_vjp_vout(), init=dict(_dx=self.cot))
        pin['grad'].copy_(g.t, non_blocking=True)
        cur.wait_stream(self.side_down)
        return dx, g

    def step_e2e(self):
        r = self.e2e_call()
        self.env.torch.cuda.current_stream().synchronize()
        return r


def timed(env, fn, steps, warmup, label, launches_per_call=None):
    """W warm-up calls, then K calls each bracketed by CUDA events on the launching stream with an
    L2 flush in between; returns (total ms as max over ranks, launches, host wall seconds)"""
    torch = env.torch
    from vmad_b200 import _capi as C
    for _ in range(warmup):
        fn()
    # the graph, plans and layout caches built during warm-up are long-lived: move them out of
    # the cyclic collector's generations so that a full collection (tens of ms of host time,
    # which lock-stepped ranks all wait for) does not land inside the timed steps
    gc.collect()
    gc.freeze()
    env.barrier()
    l0 = C.launch_count()
    per_step = []
    mallocs = [torch.cuda.memory_stats().get('num_device_alloc', 0)]
    wall0 = time.perf_counter()
    for _ in range(steps):
        env.flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        per_step.append(e0.elapsed_time(e1))
        mallocs.append(torch.cuda.memory_stats().get('num_device_alloc', 0))
    env.log("%s per-step ms: %s; cudaMalloc calls per step: %s" % (
        label, ["%.2f" % v for v in per_step], [b - a for a, b in zip(mallocs[:-1], mallocs[1:])]))
    env.barrier()
    wall = time.perf_counter() - wall0
    launches = C.launch_count() - l0
    if launches_per_call is not None:
        launches = launches_per_call * steps        # replayed launches are not re-counted
    return env.max_over_ranks(sum(per_step)), int(launches), wall


def profile_step(env, prob, ms_per_step):
    """ONE more eager step with every launch of the library timed in place (CUDA events on the
    launching stream, vpm_profile_*): per entry point (group) and per kernel -- time, share of the
    step, algorithmic bytes stated by the entry point, achieved GB/s and fraction of the peak"""
    from vmad_b200 import _capi as C
    peak, peak_src = measured_peak()
    prob.step_resident()
    env.torch.cuda.synchronize()
    with C.profile() as p:
        prob.step_resident()
        env.torch.cuda.synchronize()
    groups = collections.OrderedDict()
    kernels = collections.OrderedDict()
    total = 0.0
    for g, k, ms, b, gb in p.records:
        total += ms
        e = groups.setdefault(g, dict(calls=0, launches=0, ms=0.0, bytes=0.0))
        e['launches'] += 1
        e['ms'] += ms
        if gb > 0:
            e['calls'] += 1
            e['bytes'] += gb
        kk = kernels.setdefault((g, k), dict(launches=0, ms=0.0, group=g))
        kk['launches'] += 1
        kk['ms'] += ms
    out = []
    for g, e in sorted(groups.items(), key=lambda kv: -kv[1]['ms']):
        rec = dict(entry_point=g, calls=e['calls'], launches=e['launches'], ms=e['ms'],
                   share_of_step=e['ms'] / total if total else None,
                   avg_call_ms=e['ms'] / max(e['calls'], 1))
        if e['bytes'] > 0 and e['ms'] > 0:
            ach = e['bytes'] / (e['ms'] * 1e-3) / 1e9
            rec.update(algorithmic_bytes_per_call=e['bytes'] / max(e['calls'], 1), achieved=ach, frac=ach / peak)
        rec['kernels'] = {k[1]: dict(launches=v['launches'], ms=v['ms']) for k, v in kernels.items() if k[0] == g}
        out.append(rec)
    return dict(step_ms_profiled=total, step_ms_timed=ms_per_step, launches=len(p.records),
                entry_points=out, peak=peak, peak_source=peak_src)


def roofline_from_profile(prof, nmesh, algorithmic_step_bytes, ms_per_step):
    """the `roofline` object of the JSON line: the hand-written entry point with the largest
    measured share of the step; every entry point >= 5 % of the step is listed beside it"""
    peak = prof['peak']
    # the dominant kernel is chosen among the hand-written HBM-bound entry points: cuFFT executions are
    # library code, and the peer puts are bounded by NVLink (their bytes are stores into the peers'
    # memory: reported against 900 GB/s per direction, the flag waits for the peers included)
    own = [e for e in prof['entry_points'] if 'frac' in e and
           not e['entry_point'].startswith(('r2c', 'c2r', 'c2c', 'peer_put', 'copy2'))]
    dom = own[0] if own else None
    listed = []
    for e in prof['entry_points']:
        if e['share_of_step'] is None or e['share_of_step'] < 0.02:
            continue
        rec = dict(entry_point=e['entry_point'], share_of_step=e['share_of_step'], calls=e['calls'],
                   avg_call_ms=e['avg_call_ms'], algorithmic_bytes_per_call=e.get('algorithmic_bytes_per_call'),
                   achieved=e.get('achieved'), frac=e.get('frac'), kernels=e['kernels'],
                   traffic=recorded_traffic(nmesh, e['entry_point']), bound='hbm')
        if e['entry_point'] == 'peer_put_rows':
            # the rows that really travel are counted on the device (fixed-capacity routing); the host only
            # knows the capacities, so no bandwidth figure is derived: these launches are latency-bound
            # (flag protocol + a few MB of boundary particles)
            rec.update(bound='nvlink', achieved=None, frac=None, algorithmic_bytes_per_call=None,
                       what="boundary particles to the neighbouring slabs; bytes known on the device only")
        elif e['entry_point'].startswith('peer_put') and e.get('achieved') is not None:
            world = int(os.environ.get('WORLD_SIZE', '1'))
            stored = 0.5 * e['achieved'] * (world - 1) / max(world, 1)      # GB/s of remote stores (half of r + w)
            rec.update(bound='nvlink', achieved=stored, frac=stored / 900.0, peak=900.0,
                       what="bytes stored into the other ranks' memory / time incl. the flag waits for the peers")
        listed.append(rec)
    r = {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": prof['peak_source'],
         "kernel": None, "achieved": None, "frac": None, "traffic": None}
    if dom is not None:
        names = sorted(dom['kernels'].items(), key=lambda kv: -kv[1]['ms'])
        r.update(kernel=names[0][0], entry_point=dom['entry_point'], achieved=dom['achieved'], frac=dom['frac'],
                 traffic=recorded_traffic(nmesh, dom['entry_point']), launches=dom['calls'],
                 avg_launch_ms=dom['avg_call_ms'],
                 algorithmic_bytes_per_launch=dom['algorithmic_bytes_per_call'],
                 share_of_step=dom['share_of_step'])
    r["how"] = ("one extra eager forward+vjp step with every launch of libvmadpm.so bracketed by CUDA events on "
                "the launching stream (vpm_profile_start/_read); achieved = algorithmic bytes stated by the entry "
                "point (SURVEY.md 8d figures x the particles / cells of the call, zero fill included) / measured "
                "time; `kernel` = the hand-written entry point with the largest share of the step")
    r["entry_points"] = listed
    paint = [e for e in prof['entry_points'] if e['entry_point'] == 'paint' and 'frac' in e]
    if paint:
        r["north_star_paint"] = dict(frac=paint[0]['frac'], achieved=paint[0]['achieved'],
                                     avg_call_ms=paint[0]['avg_call_ms'], calls=paint[0]['calls'],
                                     share_of_step=paint[0]['share_of_step'],
                                     what="single-mesh CIC paint (zero fill + k_paint_*), 24 B/particle + 8 B/cell")
    r["whole_step"] = dict(algorithmic_bytes=algorithmic_step_bytes, ms=ms_per_step,
                           achieved=algorithmic_step_bytes / (ms_per_step * 1e-3) / 1e9,
                           frac=algorithmic_step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                           what="sum of the algorithmic bytes of every entry point of one step / ms_per_step")
    return r


def run_main_problem(env, args, dims, label, steps, warmup, want_profile=True, want_e2e=True, try_graph=True,
                     clocks=True):
    """build, warm up, profile, time e2e, time resident (graph replay on one GPU when it fits)"""
    torch = env.torch
    from vmad_b200 import _capi as C
    t_build = time.perf_counter()
    torch.cuda.reset_peak_memory_stats()
    prob = Problem(env, dims, args.nsteps)
    env.log("%s: problem built in %.1f s (%d local particles)" % (label, time.perf_counter() - t_build, prob.npl))
    for _ in range(2):
        prob.step_resident()
    torch.cuda.synchronize()
    res = dict()
    if env.world > 1:
        # the first steps ran with the exact routing (per-peer counts read back by the host on every
        # decompose) and recorded those counts; now switch to fixed-capacity segments with the counts
        # left on the device -- no host synchronisation inside a step, so it can be captured
        ex_ms, _, _ = timed(env, prob.step_resident, 2, 0, label + " exact routing")
        res['exact_routing_eager_ms_per_step'] = ex_ms / 2
        dx_x, g_x = prob.step_resident()
        res['routing'] = "exact (host reads the per-peer counts of every decompose)"
        if not args.no_static:
            try:
                plan = prob.pm.plan_static_routing()
                prob.step_resident()
                dx_s, g_s = prob.step_resident()
                prob.pm.check_routing()
                res['static_routing_vs_exact'] = dict(
                    gradient_rel_l2=rel_l2_across_ranks(env, g_s.t, g_x.t), dx_rel_l2=rel_l2_across_ranks(env, dx_s.t, dx_x.t),
                    what="fixed-capacity routing (counts stay on the device) vs the exact routing, same step")
                res['routing'] = ("fixed-capacity segments planned from the warm-up step (capacity matrix %s), counts "
                                  "stay on the device: no host synchronisation inside a step" % plan.cap.tolist())
                del dx_s, g_s
            except Exception as e:
                env.log("%s: static routing unavailable (%s)" % (label, str(e).splitlines()[0][:800]))
                prob.pm._route_plan = None
        del dx_x, g_x
    l0 = C.launch_count()
    prob.step_resident()
    launches_per_step = C.launch_count() - l0
    res['launches_per_step'] = int(launches_per_step)
    # eager timing first (its memory goes back to the driver before a graph takes its own pool)
    eager_ms, eager_launches, _ = timed(env, prob.step_resident, max(2, min(steps, 5)), 1, label + " eager")
    res['eager_ms_per_step'] = eager_ms / max(2, min(steps, 5))
    if want_profile:
        res['profile'] = profile_step(env, prob, res['eager_ms_per_step'])
    if want_e2e:
        e2e_ms, _, _ = timed(env, prob.step_e2e, steps, max(1, warmup // 2), label + " e2e")
        res['e2e_ms_per_step'] = e2e_ms / steps
        res['h2d'] = prob.h2d_bytes
        res['d2h'] = prob.d2h_bytes
    # keep the eager results for parity checks
    dx_e, g_e = prob.step_resident()
    torch.cuda.synchronize()
    res['eager_result'] = (dx_e, g_e)
    res['e2e_result_vs_resident'] = None
    if want_e2e:
        pin = prob.pinned
        prob.step_e2e()
        a, b = pin['grad'], g_e.t.cpu()
        c, d = pin['dx'], dx_e.t.cpu()
        res['e2e_result_vs_resident'] = dict(
            gradient_rel_l2=float((a - b).abs().pow(2).sum().sqrt() / b.abs().pow(2).sum().sqrt()),
            dx_rel_l2=float((c - d).pow(2).sum().sqrt() / d.pow(2).sum().sqrt()))
        del a, b, c, d
    res['peak_hbm_gb_eager'] = torch.cuda.max_memory_allocated() / 1e9
    sampler = ClockSampler(env.local)
    use_graph = False
    fn = prob.step_resident
    if try_graph and not args.no_graph and (env.world == 1 or prob.pm._route_plan is not None):
        from vmad_b200.cudagraph import GraphedCall
        try:
            gc.collect()
            torch.cuda.empty_cache()
            graphed = GraphedCall(prob.step_resident, warmup=1, release_memory=True)
            use_graph = True
        except Exception as e:      # typically: the graph's private pool does not fit beside the tape
            env.log("%s: CUDA graph capture failed (%s); timing the eager step" % (label, str(e).splitlines()[0][:200]))
        if env.world > 1 and int(round(env.sum_over_ranks(1.0 if use_graph else 0.0)[0])) != env.world:
            use_graph = False       # all ranks or none: a replay spins on flags only another replay raises
        if use_graph:
            def fn():
                return graphed()
        else:
            graphed = None
            gc.collect()
            torch.cuda.empty_cache()
    if clocks and env.rank == 0:
        sampler.start()
    total_ms, launches, wall = timed(env, fn, steps, warmup, label + (" graph" if use_graph else " resident"),
                                     launches_per_call=launches_per_step if use_graph else None)
    res['clocks'] = sampler.stop() if (clocks and env.rank == 0) else None
    if env.world > 1:
        prob.pm.check_routing()
    res.update(ms_per_step=total_ms / steps, launches=launches, use_graph=use_graph, host_wall_s=wall)
    if use_graph:
        dx_g, g_g = fn()
        torch.cuda.synchronize()
        res['graph_vs_eager_rel_l2'] = float((g_g.t - g_e.t).abs().pow(2).sum().sqrt() / g_e.t.abs().pow(2).sum().sqrt())
    res['peak_hbm_gb'] = torch.cuda.max_memory_allocated() / 1e9
    res['problem'] = prob
    return res


def rel_l2_across_ranks(env, a, b):
    """|| a - b || / || b || with the sums taken over all ranks (torch tensors on the device)"""
    d = (a - b)
    s = env.sum_over_ranks([float((d.abs() ** 2).sum().item()), float((b.abs() ** 2).sum().item())])
    return float((s[0] / s[1]) ** 0.5) if s[1] > 0 else float('nan')


def nrank_vs_one_rank(env, args, prob, result):
    """the main problem again on rank 0 ALONE (one GPU, no communicator) and the distributed
    displacement / gradient compared with it over all ranks"""
    torch, dist = env.torch, env.dist
    pm = prob.pm
    dims = prob.dims
    dx_d, g_d = result
    ntot = int(numpy.prod(dims))
    dx_full = torch.empty((ntot, 3), dtype=torch.float64, device='cuda')
    g_full = torch.empty((dims[0], dims[1], dims[2] // 2 + 1), dtype=torch.complex128, device='cuda')
    if env.rank == 0:
        solo = Problem(env, dims, args.nsteps, comm=None, pinned=False)
        dx1, g1 = solo.step_resident()
        dx_full.copy_(dx1.t)
        g_full.copy_(g1.t)
        del solo, dx1, g1
        gc.collect()
        torch.cuda.empty_cache()
    dist.broadcast(dx_full, src=0)
    gr = torch.view_as_real(g_full)
    dist.broadcast(gr, src=0)
    npl = prob.npl
    out = dict(dx_rel_l2=rel_l2_across_ranks(env, dx_d.t, dx_full[env.rank * npl:(env.rank + 1) * npl]),
               gradient_rel_l2=rel_l2_across_ranks(env, g_d.t, g_full[:, pm.cstart1:pm.cstart1 + pm.cshape[1]]),
               what="%d-rank slab-decomposed result of the main problem vs the same problem on one GPU "
                    "(rank 0, no communicator), norms over all ranks" % env.world)
    del dx_full, g_full
    return out


def small_cross_rank_checks(env):
    """P-rank == 1-rank on small problems the driver can see: chi-square objective / gradient /
    Gauss-Newton product (contrib/chisquare.py:69-93) and FastPMSimulation(B=2) forward + vjp"""
    import torch
    from vmad_b200 import autooperator, Builder
    from vmad_b200.contrib.chisquare import MPIChiSquareProblem
    from vmad_b200.lib import fastpm, linalg
    from vmad_b200.pm import DeviceArray, ParticleMesh, SelfComm
    N, L = 32, 100.0
    stages = numpy.linspace(0.1, 1.0, 4)
    rng = numpy.random.default_rng(5)
    data_h = 1.0 + 0.3 * rng.standard_normal((N, N, N))
    wn_h = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    v_h = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    cot_h = rng.standard_normal((N ** 3, 3))
    out = {}

    def run(comm):
        pm = ParticleMesh([N] * 3, L, comm=comm)
        q = pm.generate_uniform_particle_grid(shift=0.0)
        data = pm.create('real', value=data_h)

        @autooperator('x->model')
        def forward(x):
            rhok = fastpm.induce_correlation(x, power, pm)
            dx, p, f = fastpm.nbody(rhok, q, stages, Cosmology, pm)
            xf = linalg.add(dx, q)
            layout = fastpm.decompose(xf, pm)
            return dict(model=fastpm.paint(xf, 1.0, layout, pm))

        @autooperator('model->r')
        def residual(model):
            return dict(r=(model - data) * 0.5)

        problem = MPIChiSquareProblem(pm.comm, forward, [residual])
        s = pm.create('complex', value=wn_h)
        v = pm.create('complex', value=v_h)
        r = dict(objective=float(problem.objective(s)), gradient=problem.gradient(s).t.clone(),
                 gn=problem.hessian_vector_product(s, v).t.clone())
        sim = fastpm.FastPMSimulation(stages, Cosmology, pm, B=2, q=q)
        with Builder() as m:
            x = m.input('x')
            dx, p, f = sim.run(fastpm.induce_correlation(x, power, pm))
            m.output(dx=dx, p=p, f=f)
        npl = len(q)
        lo = pm.rank * npl
        (dx,), (g,) = m.compute_with_vjp(init=dict(x=s), v=dict(_dx=DeviceArray.from_host(cot_h[lo:lo + npl])))
        r.update(sim_dx=dx.t.clone(), sim_vjp=g.t.clone())
        return pm, r

    pm, dist_r = run(env.comm)
    ref = None
    if env.rank == 0:
        _, ref = run(SelfComm())
    shapes = dict(gradient=(N, N, N // 2 + 1), gn=(N, N, N // 2 + 1), sim_vjp=(N, N, N // 2 + 1))
    obj = torch.zeros(1, dtype=torch.float64, device='cuda')
    if env.rank == 0:
        obj[0] = ref['objective']
    env.dist.broadcast(obj, src=0)
    out['chisquare_objective_rel'] = abs(dist_r['objective'] - float(obj.item())) / abs(float(obj.item()))
    cs = slice(pm.cstart1, pm.cstart1 + pm.cshape[1])
    for key in ('gradient', 'gn', 'sim_vjp'):
        full = torch.empty(shapes[key], dtype=torch.complex128, device='cuda')
        if env.rank == 0:
            full.copy_(ref[key])
        env.dist.broadcast(torch.view_as_real(full), src=0)
        out[{'gradient': 'chisquare_gradient_rel_l2', 'gn': 'chisquare_gauss_newton_rel_l2',
             'sim_vjp': 'fastpmsimulation_B2_vjp_rel_l2'}[key]] = rel_l2_across_ranks(env, dist_r[key], full[:, cs])
    full = torch.empty((N ** 3, 3), dtype=torch.float64, device='cuda')
    if env.rank == 0:
        full.copy_(ref['sim_dx'])
    env.dist.broadcast(full, src=0)
    npl = N ** 3 // env.world
    out['fastpmsimulation_B2_dx_rel_l2'] = rel_l2_across_ranks(env, dist_r['sim_dx'], full[env.rank * npl:(env.rank + 1) * npl])
    out['what'] = ("%d ranks vs one GPU (rank 0, SelfComm): MPIChiSquareProblem objective / gradient / Gauss-Newton "
                   "product over a 3-step nbody + paint at 32^3, and FastPMSimulation(B=2).run forward + vjp" % env.world)
    return out


def gn_step_record(env, args, nmesh, nsteps=5, repeats=2):
    """BASELINE.json configs[3]: one Gauss-Newton Hessian-vector product (jvp then vjp through the
    `precompute` replay, chisquare.py:76-93, model.py:262-311) of a chi-square over a 5-step FastPM
    forward model at nmesh^3, slab-decomposed over the ranks"""
    import torch
    from vmad_b200 import autooperator
    from vmad_b200.contrib.chisquare import MPIChiSquareProblem
    from vmad_b200.lib import fastpm, linalg
    from vmad_b200.pm import ParticleMesh, RealField
    dims = [nmesh] * 3
    torch.cuda.reset_peak_memory_stats()
    pm = ParticleMesh(dims, [100.0 * n / 32 for n in dims], comm=env.comm)
    pm.force_recompute = int(numpy.prod(pm.rshape)) >= 256 ** 3
    q = pm.generate_uniform_particle_grid(shift=0.0)
    stages = numpy.linspace(0.1, 1.0, nsteps + 1)
    n0l = pm.rshape[0]
    mk = lambda seed: RealField(pm, torch.from_numpy(noise_planes(dims, pm.rstart0, n0l, seed)).cuda())
    s = pm._r2c(mk(SEED), pm.Ntot ** -0.5)
    v = pm._r2c(mk(SEED + 1), pm.Ntot ** -0.5)
    data = mk(SEED + 2) * 0.3 + 1.0

    @autooperator('x->model')
    def forward(x):
        rhok = fastpm.induce_correlation(x, power, pm)
        dx, p, f = fastpm.nbody(rhok, q, stages, Cosmology, pm)
        xf = linalg.add(dx, q)
        layout = fastpm.decompose(xf, pm)
        return dict(model=fastpm.paint(xf, 1.0, layout, pm))

    @autooperator('model->r')
    def residual(model):
        return dict(r=(model - data) * 0.5)

    problem = MPIChiSquareProblem(pm.comm, forward, [residual])
    ts = []
    hv = None
    for i in range(repeats + 1):
        env.barrier()
        env.flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        hv = problem.hessian_vector_product(s, v)
        e1.record()
        e1.synchronize()
        if i > 0:
            ts.append(e0.elapsed_time(e1))
    ms = env.max_over_ranks(float(numpy.mean(ts)))
    # size-independent property: the Gauss-Newton operator is symmetric positive semi-definite,
    # <v, H v> = 2 |J v|^2 >= 0, and linear: H(2 v) = 2 H v
    vHv = pm.comm.allreduce(float(torch.view_as_real(hv.t * v.t.conj())[..., 0].sum().item())) if env.world > 1 \
        else float(torch.view_as_real(hv.t * v.t.conj())[..., 0].sum().item())
    hv2 = problem.hessian_vector_product(s, v * 2.0)
    lin = rel_l2_across_ranks(env, hv2.t, hv.t * 2.0)
    rec = dict(metric="Gauss-Newton Hessian-vector products/s", value=1e3 / ms, unit="products/s", ms_per_product=ms,
               config=dict(workload="one MPIChiSquareProblem.hessian_vector_product (forward + precompute tape + jvp + vjp) "
                                    "over a %d-step FastPM forward model + paint, %d^3 particles / mesh, fp64" % (nsteps, nmesh),
                           nmesh=nmesh, nsteps=nsteps, n_gpus=env.world),
               properties=dict(v_dot_Hv=vHv, linearity_rel_l2=lin,
                               what="<v, H v> >= 0 (H = 2 J^T J) and H(2 v) = 2 H v to rounding"),
               peak_hbm_gb=torch.cuda.max_memory_allocated() / 1e9)
    return rec


def fastpm_large_record(env, args, nmesh, steps=2):
    """BASELINE.json configs[4]: the 10-step forward+vjp at nmesh^3 (1024^3 on 8 GPUs) in the
    memory-lean mode; size-independent property instead of an oracle: the directional derivative
    along the cotangent direction from the vjp equals the one from a forward difference is too
    expensive here, so the check is linearity of the vjp in the cotangent"""
    import torch
    dims = [nmesh] * 3
    torch.cuda.reset_peak_memory_stats()
    prob = Problem(env, dims, args.nsteps, force_recompute=True, pinned=False)
    prob.step_resident()
    ms, _, _ = timed(env, prob.step_resident, steps, 1, "fastpm_%d" % nmesh)
    dx, g = prob.step_resident()
    g1 = g.t.clone()
    prob.cot = prob.cot * 2.0
    _, g2 = prob.step_resident()
    lin = rel_l2_across_ranks(env, g2.t, g1 * 2.0)
    mom = env.sum_over_ranks([float(dx.t.sum().item()), float(dx.t.abs().sum().item())])
    rec = dict(metric="FastPM forward+vjp steps/s at %d^3" % nmesh, value=args.nsteps / (ms / steps * 1e-3), unit=UNIT,
               ms_per_step=ms / steps, steps=steps,
               config=dict(workload=workload_name(nmesh, args.nsteps), nmesh=nmesh, nsteps=args.nsteps,
                           n_gpus=env.world, mode="memory-lean (force_recompute)"),
               properties=dict(vjp_linearity_rel_l2=lin, mean_displacement_over_mean_abs=float(mom[0] / mom[1]),
                               what="vjp(2 c) = 2 vjp(c) to rounding; the mean displacement vanishes (momentum conservation of the PM force)"),
               peak_hbm_gb=torch.cuda.max_memory_allocated() / 1e9)
    del prob
    return rec


def run_gpu(args):
    env = Env()
    torch = env.torch
    world, rank = env.world, env.rank
    N = args.nmesh
    if N % world:
        raise SystemExit("bench.py: --nmesh must be divisible by the number of GPUs")
    t_start = time.perf_counter()

    # ---- main record: the metric's configuration, ONE global problem ------------------------------
    main = run_main_problem(env, args, [N] * 3, "main %d^3" % N, args.steps, args.warmup)
    prob = main.pop('problem')
    ms_per_step = main['ms_per_step']
    value = args.nsteps / (ms_per_step * 1e-3)
    e2e_value = args.nsteps / (main['e2e_ms_per_step'] * 1e-3)
    prof = main.pop('profile')
    alg_step_bytes = sum(e.get('algorithmic_bytes_per_call', 0.0) * e['calls'] for e in prof['entry_points'])
    roofline = roofline_from_profile(prof, N, alg_step_bytes, ms_per_step)
    parity = {}
    side = {}
    if main.get('e2e_result_vs_resident'):
        parity['e2e_host_results_vs_resident'] = main['e2e_result_vs_resident']
    if main.get('graph_vs_eager_rel_l2') is not None:
        parity['graph_replay_vs_eager_gradient_rel_l2'] = main['graph_vs_eager_rel_l2']
    if main.get('static_routing_vs_exact') is not None:
        parity['static_routing_vs_exact'] = main['static_routing_vs_exact']
    eager_result = main.pop('eager_result')
    if world > 1 and not args.no_side:
        try:
            # free the tape-sized pools of the main problem before rank 0 runs it alone
            dx_d, g_d = eager_result
            del prob.model
            gc.collect()
            torch.cuda.empty_cache()
            parity['nrank_vs_1rank_main'] = nrank_vs_one_rank(env, args, prob, (dx_d, g_d))
        except Exception as e:
            parity['nrank_vs_1rank_main'] = dict(error=str(e).splitlines()[0][:300])
            env.log("nrank_vs_1rank_main failed:", e)
    peer = prob.pm._peer() is not None if world > 1 else None
    del prob, eager_result
    gc.collect()
    torch.cuda.empty_cache()
    env.log("main record done at %.0f s" % (time.perf_counter() - t_start))

    # ---- side record: 128^3 cells and particles per GPU (round-1 configuration) ---------------------
    if not args.no_side:
        dims = weak_shape(args.side_nmesh, world)
        w = run_main_problem(env, args, dims, "side %s" % dims, args.steps, args.warmup, want_profile=True,
                             clocks=False)
        wp = w.pop('problem')
        wprof = w.pop('profile')
        wdx, wg = w.pop('eager_result')
        side['weak_%d' % args.side_nmesh] = dict(
            metric="FastPM forward+vjp steps/s, %d^3 cells and particles per GPU (weak scaling; value = n_gpus x nsteps / s)" % args.side_nmesh,
            value=world * args.nsteps / (w['ms_per_step'] * 1e-3), unit=UNIT, ms_per_step=w['ms_per_step'],
            e2e_value=world * args.nsteps / (w['e2e_ms_per_step'] * 1e-3), eager_ms_per_step=w['eager_ms_per_step'],
            global_mesh=dims, launches_per_step=w['launches_per_step'], cuda_graph=w['use_graph'],
            routing=w.get('routing'), exact_routing_eager_ms_per_step=w.get('exact_routing_eager_ms_per_step'),
            static_routing_vs_exact=w.get('static_routing_vs_exact'),
            graph_vs_eager_rel_l2=w.get('graph_vs_eager_rel_l2'),
            roofline=roofline_from_profile(wprof, args.side_nmesh, sum(e.get('algorithmic_bytes_per_call', 0.0) * e['calls'] for e in wprof['entry_points']), w['ms_per_step']))
        env.log("side record done at %.0f s" % (time.perf_counter() - t_start))
        # ---- parity against the oracle: the same 128^3 x nsteps problem on the CPU port -----------------
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            nm = args.side_nmesh
            ts, (odx, og) = cpu_fwd_vjp(nm, args.nsteps, threads, repeats=1, keep_result=True)
            sec = ts[0]
            gdx, gg = wdx.t.cpu().numpy(), wg.t.cpu().numpy()
            parity['vs_oracle_%d' % nm] = dict(
                gradient_rel_l2=float(numpy.linalg.norm(gg - og) / numpy.linalg.norm(og)),
                dx_rel_l2=float(numpy.linalg.norm(gdx - odx) / numpy.linalg.norm(odx)), bar=1e-10,
                what="GPU (eager step of the side record) vs the numpy/C oracle, the same %d^3 x %d-step forward+vjp on the same inputs"
                     % (nm, args.nsteps))
            scale = (float(nm) / N) ** 3
            cpu_baseline = {"value": args.nsteps / sec * scale, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": cpu_sample_text(args.nsteps, nm, N, threads, sec),
                            "measured_ms_per_sample": sec * 1e3, "sample_nmesh": nm,
                            "value_at_sample_size": args.nsteps / sec}
        del wp, wdx, wg
        gc.collect()
        torch.cuda.empty_cache()
    else:
        cpu_baseline = None
    if world > 1 and not args.no_side:
        try:
            parity['p_rank_vs_1_rank_small'] = small_cross_rank_checks(env)
        except Exception as e:
            parity['p_rank_vs_1_rank_small'] = dict(error=str(e).splitlines()[0][:300])
            env.log("small cross-rank checks failed:", e)
        gc.collect()
        torch.cuda.empty_cache()
    if world >= 8 and not args.no_side and not args.no_large:
        for name, fn in (('gn_step_512', lambda: gn_step_record(env, args, 512)),
                         ('fastpm_1024', lambda: fastpm_large_record(env, args, 1024))):
            try:
                side[name] = fn()
            except Exception as e:
                side[name] = dict(error=str(e).splitlines()[0][:300])
                env.log("%s failed:" % name, e)
            gc.collect()
            torch.cuda.empty_cache()
            env.log("%s done at %.0f s" % (name, time.perf_counter() - t_start))

    if rank == 0:
        bad = []
        for k, v in parity.items():
            if isinstance(v, dict):
                for kk, vv in v.items():
                    if kk.endswith(('rel_l2', '_rel')) and not (vv < 1e-10):
                        bad.append((k, kk, vv))
        gvals = [v for k, d in parity.items() if isinstance(d, dict) for kk, v in d.items() if kk == 'gradient_rel_l2']
        line = {
            "metric": METRIC_FMT % args.nmesh, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(N, args.nsteps),
            "run": {"parallelism": ("slab decomposition of the ONE global mesh over %d GPUs (mesh axis 0; %s for particle "
                                    "exchange and FFT transposes)" % (world, "peer-memory stores over NVLink (CUDA IPC)" if peer
                                                                      else "NCCL all-to-all")) if world > 1 else "1 gpu",
                    "l2": "256 MiB flush write between timed steps",
                    "step": "one Model.compute_with_vjp = %d FastPM steps forward + reverse" % args.nsteps,
                    "launch": ("CUDA graph replay of the captured compute_with_vjp (same input buffers; %d kernel "
                               "launches per replay)" % main['launches_per_step']) if main['use_graph']
                    else "eager (python VM + ctypes, one launch per kernel)",
                    "eager_ms_per_step": main['eager_ms_per_step'], "launches_per_step": main['launches_per_step'],
                    "routing": main.get('routing'), "exact_routing_eager_ms_per_step": main.get('exact_routing_eager_ms_per_step'),
                    "peak_hbm_gb": main['peak_hbm_gb'], "wall_s_total": time.perf_counter() - t_start},
            "clocks": main['clocks'],
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": main['h2d'], "d2h_bytes_per_step": main['d2h'],
                    "what": "host white noise + cotangent in, displacement dx + gradient out (pinned buffers), per rank"},
            "gpu_launches": int(main['launches']),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "parity_rel_l2": max(gvals) if gvals else None,
            "parity_failures": bad,
            "side": side,
        }
        emit(line)
    if world > 1:
        env.dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# microbenchmark (BASELINE.json configs[2])
# ---------------------------------------------------------------------------------------------
def run_micro(args):
    """CIC paint / readout apl + vjp + jvp at 256^3 with 2^24 .. 2^27 particles, per GPU: every
    rank runs the same kernels on its own particles (the path shards without an exchange once the
    particles are routed; routing is timed separately as decompose / exchange); L2 flushed between
    launches, clocks sampled, CUDA events on the launching stream"""
    env = Env()
    torch = env.torch
    from vmad_b200.pm import DeviceArray, ParticleMesh
    peak, peak_src = measured_peak()
    N = args.micro_nmesh
    L = 100.0 * N / 32
    pm = ParticleMesh([N] * 3, L)
    Nc = N ** 3
    gen = torch.Generator(device='cuda').manual_seed(1 + env.rank)
    sampler = ClockSampler(env.local)
    if env.rank == 0:
        sampler.start()
    rows = []

    def t_of(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            env.flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return float(numpy.median(ts))

    for dist_name in args.micro_dist:
        for l2 in args.log2np:
            n = 1 << l2
            if dist_name == 'uniform':
                x = torch.rand((n, 3), generator=gen, device='cuda', dtype=torch.float64) * L
            else:   # FastPM-like: lattice + Gaussian displacement of 1.5 cells
                side_n = round(n ** (1 / 3))
                n = side_n ** 3
                ax = torch.arange(side_n, device='cuda', dtype=torch.float64) * (L / side_n)
                x = torch.stack(torch.meshgrid(ax, ax, ax, indexing='ij'), -1).reshape(-1, 3)
                x = x + torch.randn((n, 3), generator=gen, device='cuda', dtype=torch.float64) * (1.5 * L / N)
            x = DeviceArray(x)
            m = DeviceArray(torch.rand((n,), generator=gen, device='cuda', dtype=torch.float64) + 0.5)
            f = pm.create('real')
            f.t.copy_(torch.randn((N, N, N), generator=gen, device='cuda', dtype=torch.float64))
            lay = pm.decompose(x)
            xs = lay.exchange(x)
            ms_ = lay.exchange(m)
            v3 = DeviceArray(torch.randn((n, 3), generator=gen, device='cuda', dtype=torch.float64))
            v3s = lay.exchange(v3)
            ops = [
                ('decompose (cell sort)', lambda: pm.decompose(x), 48.0 * n),
                ('exchange [n,3]', lambda: lay.exchange(x), 52.0 * n),
                ('gather [n,3]', lambda: lay.gather(xs), 52.0 * n),
                ('paint apl', lambda: pm.paint(xs, 1.0), 24.0 * n + 8.0 * Nc),
                ('paint apl (mass array)', lambda: pm.paint(xs, ms_), 32.0 * n + 8.0 * Nc),
                ('paint vjp', lambda: pm.paint_vjp(f, xs, mass=1.0), 56.0 * n + 8.0 * Nc),
                ('paint jvp', lambda: pm.paint_jvp(xs, mass=1.0, v_pos=v3s), 48.0 * n + 8.0 * Nc),
                ('readout apl', lambda: f.readout(xs), 32.0 * n + 8.0 * Nc),
                ('readout vjp', lambda: f.readout_vjp(xs, ms_), 72.0 * n + 16.0 * Nc),
                ('readout jvp', lambda: f.readout_jvp(xs, v_self=f, v_pos=v3s), 56.0 * n + 16.0 * Nc),
                ('paint x3 (force vjp)', lambda: pm.paint_cols3(xs, v3s), 48.0 * n + 24.0 * Nc),
                ('readout x3 (force apl)', lambda: pm.readout3(f, f, f, xs), 48.0 * n + 24.0 * Nc),
            ]
            for name, fn, nbytes in ops:
                t = t_of(fn, args.micro_iters)
                tmax = env.max_over_ranks(t)
                rows.append(dict(op=name, dist=dist_name, np_per_gpu=n, ms=tmax * 1e3,
                                 gparticles_s=env.world * n / tmax / 1e9,
                                 gparticles_s_per_gpu=n / tmax / 1e9,
                                 achieved_gbs_per_gpu=nbytes / tmax / 1e9, frac=nbytes / tmax / 1e9 / peak))
            del x, m, f, lay, xs, ms_, v3, v3s
            gc.collect()
            torch.cuda.empty_cache()
    clocks = sampler.stop() if env.rank == 0 else None
    if env.rank == 0:
        emit({"metric": "CIC paint/readout Gparticles/s", "unit": "Gparticles/s", "n_gpus": env.world,
              "config": {"workload": "CIC paint/readout apl+vjp+jvp microbench, %d^3 mesh per GPU, 2^%s particles per GPU"
                                     % (N, ",".join(str(v) for v in args.log2np)),
                         "l2": "256 MiB flush write before every timed launch", "iters": args.micro_iters,
                         "timing": "CUDA events on the launching stream, median, max over ranks"},
              "peak": peak, "peak_source": peak_src, "clocks": clocks, "rows": rows})
    if env.world > 1:
        env.dist.destroy_process_group()


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
    ap.add_argument('--nmesh', type=int, default=512, help='the global problem: nmesh^3 particles and cells')
    ap.add_argument('--nsteps', type=int, default=10)
    ap.add_argument('--side-nmesh', type=int, default=128, help='cells / particles per GPU of the weak-scaling side record')
    ap.add_argument('--cpu-nmesh', type=int, default=128,
                    help='largest mesh the CPU arm / cpu_baseline leg will run (bounded sample)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='run every step eagerly through the VM')
    ap.add_argument('--no-static', action='store_true', help='keep the exact (host-synchronised) particle routing across ranks')
    ap.add_argument('--no-side', action='store_true', help='main record only')
    ap.add_argument('--no-large', action='store_true', help='skip the 8-GPU configs[3] / configs[4] side records')
    ap.add_argument('--micro', action='store_true', help='CIC paint / readout microbenchmark instead')
    ap.add_argument('--micro-nmesh', type=int, default=256)
    ap.add_argument('--log2np', type=int, nargs='+', default=[24, 25, 26, 27])
    ap.add_argument('--micro-dist', nargs='+', default=['uniform', 'lattice'])
    ap.add_argument('--micro-iters', type=int, default=10)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.micro:
        run_micro(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
