#!/usr/bin/env python
"""bench.py -- R2C fftn+ifftn round trip (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's CPU algorithm (oracle port), host cores

A step is one fftn + one ifftn of the workload (default: slab.R2C 1024^3 double -- the north-star
configuration; it fits one GPU).  `value` is GFLOP/s by the 5*M*log2(M) convention (M = global
real points) with device-resident arrays; `e2e` is the same through the numpy API with pinned host
buffers, H2D and D2H inside the timed region.  One JSON line is printed by rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, N, precision, dealias, kwargs)
    "slab1024_f64": ("slab", (1024, 1024, 1024), "double", None, {}),
    "slab1024_f64_32": ("slab", (1024, 1024, 1024), "double", "3/2-rule", {}),
    "slab512_f64": ("slab", (512, 512, 512), "double", None, {}),
    "slab256_f32": ("slab", (256, 256, 256), "single", None, {}),
    "slab2048z_f32": ("slab", (256, 512, 2048), "single", None, {}),
    "pencilX512_f64": ("pencil", (512, 512, 512), "double", None, dict(alignment="X", P1=2, communication="Alltoallw")),
    "pencilX1024_f64": ("pencil", (1024, 1024, 1024), "double", None, dict(alignment="X", P1=None, communication="Alltoallw")),
    "pencilY2048_f32": ("pencil", (2048, 2048, 2048), "single", None, dict(alignment="Y", P1=None, communication="Alltoallw")),
    "line4096_f32": ("line", (4096, 4096), "single", None, {}),
    "line4096_f64": ("line", (4096, 4096), "double", None, {}),
    "line2048_f64": ("line", (2048, 2048), "double", None, {}),
    "line8192_f32": ("line", (8192, 8192), "single", None, {}),
    "line16384_f32": ("line", (16384, 16384), "single", None, {}),
}
METRIC = "R2C fftn+ifftn round-trip GFLOP/s (5*M*log2(M) per transform)"
UNIT = "GFLOP/s"


def flops_roundtrip(shape):
    M = float(np.prod([float(s) for s in shape]))
    return 2.0 * 5.0 * M * math.log2(M)


def describe(name, P):
    kind, N, prec, dealias, kw = WORKLOADS[name]
    return {"workload": "%s.R2C N=%s %s fftn+ifftn round trip, dealias=%s" % (kind, "x".join(map(str, N)), prec, dealias),
            "name": name, "N": list(N), "precision": prec, "dealias": dealias,
            "decomposition": ("%s P=%d" % (kind, P)) + ("".join(" %s=%s" % kv for kv in sorted(kw.items()))),
            "l2": "operands >> 126 MB L2 (no flush needed)" if np.prod(N) * 4 > 1 << 30 else "L2 flushed between steps"}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": (float(np.median(sm)) if sm else None), "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path on this box's host cores.
#   kind "reference": the UNMODIFIED mpiFFT4py (pip-installed into baseline/_ref)
#       driven through its public API by oracle/refshim -- its MPI ranks run as threads of this
#       process (fake mpi4py, memcpy collectives) with its numpy.fft backend; pyfftw / mpi4py /
#       mpirun do not exist in this image.
#   kind "port": oracle/ (numpy restatement, pocketfft through scipy.fft workers=cores) when the
#       install is absent.
# Both are test infrastructure: nothing here is on the product path.
# ------------------------------------------------------------------------------------------------
def cpu_sample(name):
    """Bounded sample of the workload: same class / precision / dealias, smaller mesh."""
    kind, N, prec, dealias, kw = WORKLOADS[name]
    if kind == "line":
        return (4096, 4096)
    return (256, 256, 256)


def _pow2_floor(x):
    p = 1
    while 2 * p <= x:
        p *= 2
    return p


def _load_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    import load_reference
    if not load_reference.reference_available():
        return None
    try:
        load_reference.load()
    except Exception:  # noqa: BLE001
        return None
    return load_reference


def reference_roundtrip(lr, kind, N, prec, dealias, kw, P, reps, warmup=1):
    """Per-step seconds (max over ranks; `reps` timed fftn+ifftn round trips after `warmup` untimed ones) of the
    unmodified reference; the transform object and its arrays are built once."""
    from mpi4py import MPI  # the refshim's in-process stand-in

    def body():
        Nn = np.array(N, dtype=int)
        L = np.array([2 * np.pi] * len(N))
        comm = MPI.COMM_WORLD
        if kind == "slab":
            from mpiFFT4py.slab import R2C
            F = R2C(Nn, L, comm, prec)
        elif kind == "pencil":
            from mpiFFT4py.pencil import R2C
            F = R2C(Nn, L, comm, prec, **kw)
        else:
            from mpiFFT4py.line import R2C
            F = R2C(Nn, L, comm, prec)
        fwd, inv = (F.fft2, F.ifft2) if kind == "line" else (F.fftn, F.ifftn)
        rshape = F.real_shape_padded() if dealias == "3/2-rule" else F.real_shape()
        u = np.random.default_rng(1234 + comm.Get_rank()).random(rshape).astype(F.float)
        fu = np.zeros(F.complex_shape(), dtype=F.complex)
        u2 = np.zeros_like(u)
        for _ in range(max(1, warmup)):  # work arrays, subarray types
            fwd(u, fu, dealias)
            inv(fu, u2, dealias)
        ts = []
        for _ in range(reps):
            comm.barrier()
            t0 = time.perf_counter()
            fwd(u, fu, dealias)
            inv(fu, u2, dealias)
            comm.barrier()
            ts.append(time.perf_counter() - t0)
        return ts

    per_rank = lr.run_ranks(P, body)
    return [max(t) for t in zip(*per_rank)]


def port_roundtrip(kind, N, prec, dealias, workers):
    import oracle
    oracle.common.set_workers(workers)
    rt, ct = oracle.common.dtypes(prec)
    rng = np.random.default_rng(1234)
    mod = oracle.line if kind == "line" else oracle.slab
    g = mod.Geometry(N, 1)
    shape = g.real_shape_padded() if dealias == "3/2-rule" else g.real_shape()
    u = [rng.random(shape).astype(rt)]
    fwd, inv = (mod.fft2, mod.ifft2) if kind == "line" else (mod.fftn, mod.ifftn)
    t0 = time.perf_counter()
    fu = fwd(u, N, 1, dealias=dealias, precision=prec)
    inv(fu, N, 1, dealias=dealias, precision=prec)
    return time.perf_counter() - t0


def reference_ranks(name, Ns):
    """Rank count of the reference arm: as many thread-ranks as the host has cores, within what the class accepts."""
    kind = WORKLOADS[name][0]
    cores = os.cpu_count() or 1
    if kind == "pencil":
        return 8 if cores >= 8 else 4
    return max(1, min(_pow2_floor(cores), 32, Ns[0] // 2))


def cpu_time(name, Ns, reps, warmup=1):
    """(list of `reps` seconds per round trip, descriptor) of the CPU reference on mesh Ns."""
    kind, N, prec, dealias, kw = WORKLOADS[name]
    cores = os.cpu_count() or 1
    lr = _load_reference()
    if lr is not None:
        P = reference_ranks(name, Ns)
        ts = reference_roundtrip(lr, kind, Ns, prec, dealias, kw, P, reps, warmup)
        return ts, {"kind": "reference", "cores": P, "ranks": P,
                    "how": "unmodified mpiFFT4py %s.R2C (baseline/_ref) under oracle/refshim: %d ranks as threads, "
                           "numpy.fft backend, memcpy collectives" % (kind, P)}
    port_roundtrip(kind, tuple(max(32, n // 4) for n in Ns), prec, dealias, cores)  # warm-up
    ts = [port_roundtrip(kind, Ns, prec, dealias, cores) for _ in range(reps)]
    return ts, {"kind": "port", "cores": cores, "ranks": 1,
                "how": "oracle port of the reference algorithm (%s P=1, pocketfft via scipy.fft workers=%d)" % (kind, cores)}


def global_real_shape(Ns, dealias):
    return tuple(int(1.5 * n) for n in Ns) if dealias == "3/2-rule" else tuple(Ns)


def cpu_baseline(name):
    """About 10-30 s of host work on a bounded sample of the workload."""
    kind, N, prec, dealias, kw = WORKLOADS[name]
    Ns = cpu_sample(name)
    ts, d = cpu_time(name, Ns, 1)
    if kind != "line" and min(ts) < 1.5 and Ns[0] < N[0]:
        Ns = tuple(2 * n for n in Ns)
        ts, d = cpu_time(name, Ns, 2)
    t = min(ts)
    return {"value": flops_roundtrip(global_real_shape(Ns, dealias)) / t / 1e9, "unit": UNIT, "cores": d["cores"],
            "kind": d["kind"], "sample": "%s; %s %s round trip (dealias=%s), %.2f s" % (d["how"], "x".join(map(str, Ns)), prec, dealias, t),
            "seconds": t, "N": list(Ns), "host_cores": os.cpu_count()}


def _host_bytes_available():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:  # noqa: BLE001
        return 0


def reference_mesh(name, steps, warmup, budget_s=420.0):
    """The mesh the reference arm is timed on: the workload's own if the host can hold it and `steps + warmup`
    round trips fit the time budget, else the largest halving that does.  A probe on a small mesh gives the rate."""
    kind, N, prec, dealias, kw = WORKLOADS[name]
    Ns = cpu_sample(name)
    ts, d = cpu_time(name, Ns, 1)
    rate = flops_roundtrip(global_real_shape(Ns, dealias)) / min(ts)
    word = 8 if prec == "double" else 4
    cand = tuple(N)
    while True:
        pts = float(np.prod(global_real_shape(cand, dealias)))
        need = pts * word * 8  # u, fu, u2, the class's work arrays and transients of numpy.fft
        est = flops_roundtrip(global_real_shape(cand, dealias)) / rate * (steps + warmup) * 1.3 + pts * 2e-8
        if (need < 0.6 * _host_bytes_available() and est < budget_s) or cand[0] <= Ns[0]:
            return cand
        cand = tuple(n // 2 for n in cand)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    kind, N, prec, dealias, kw = WORKLOADS[name]
    Ns = reference_mesh(name, args.steps, args.warmup)
    ts, d = cpu_time(name, Ns, args.steps, max(1, args.warmup))
    dt = float(np.mean(ts))
    val = flops_roundtrip(global_real_shape(Ns, dealias)) / dt / 1e9
    # config describes what was TIMED: the mesh of this run and the reference's rank count on the host cores
    cfg = describe(name, d["ranks"])
    cfg["N"] = list(Ns)
    cfg["workload"] = "%s.R2C N=%s %s fftn+ifftn round trip, dealias=%s" % (kind, "x".join(map(str, Ns)), prec, dealias)
    cfg["decomposition"] = "%s P=%d CPU thread-ranks" % (kind, d["ranks"])
    cfg["l2"] = "host run"
    full = tuple(Ns) == tuple(N)
    if not full:
        cfg["sampled_from"] = {"N": list(N), "why": "host memory / time budget", "compare_as": "GFLOP/s rate, not ms_per_step"}
    sample = "%s on %s %s %s (%d timed round trips, mean)" % (
        d["how"], "the full mesh" if full else "a bounded sample of the %s workload:" % "x".join(map(str, N)),
        "x".join(map(str, Ns)), prec, args.steps)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64" if prec == "double" else "f32", "data": "synthetic", "config": cfg,
           "same_mesh_as_b200_arm": full,
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": d["cores"], "kind": d["kind"], "sample": sample,
                            "host_cores": os.cpu_count()},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def make_transform(m, comm, name):
    kind, N, prec, dealias, kw = WORKLOADS[name]
    Nn = np.array(N, dtype=int)
    L = np.array([2 * np.pi] * len(N))
    if kind == "slab":
        return m.Slab_R2C(Nn, L, comm, prec)
    if kind == "pencil":
        return m.Pencil_R2C(Nn, L, comm, prec, **kw)
    return m.Line_R2C(Nn, L, comm, prec)


def separable_vectors(N, dealias):
    pad = 1.5 if dealias == "3/2-rule" else 1
    rng = np.random.default_rng(4321)
    return [rng.random(int(pad * n)) + 0.25 for n in N]


def separable_spectra(vec, N, dealias, single):
    """1D spectra whose outer product is the transform of the separable field (see forward_parity)."""
    padded = dealias == "3/2-rule"
    pad = 1.5 if padded else 1
    out = []
    for ax, (v, n) in enumerate(zip(vec, N)):
        v = v.astype(np.float32).astype(np.float64) if single else v  # the real array holds v rounded to its precision
        last = ax == len(N) - 1
        f = np.fft.rfft(v) if last else np.fft.fft(v)
        if padded:
            h = n // 2
            if last:
                f = f[:h + 1].copy()
            else:
                t = np.zeros(n, dtype=np.complex128)
                t[:h + 1] = f[:h + 1]
                t[h:] += f[-h:]
                f = t
            f = f / pad
        out.append(f)
    return out


def forward_parity(F, kind, N, dealias, fwd, u, fu, torch):
    """rel. L2 error of THIS rank's forward result against the closed form for a separable field
    u = a(x) b(y) c(z): its transform is the outer product of the three 1D transforms (for the 3/2-rule: each
    padded-length 1D spectrum truncated with the Nyquist fold of slab.py:529-533 and divided by padsize),
    cut by complex_local_slice().  One extra transform; works at every rank count and decomposition, so the
    scaling runs carry forward parity for P > 1 (a round trip alone would also close under a self-inverse
    permutation mistake).  numpy's 1D FFTs of three vectors are the reference here, nothing from oracle/."""
    padded = dealias == "3/2-rule"
    pad = 1.5 if padded else 1
    dims = len(N)
    vec = separable_vectors(N, dealias)
    rs = F.real_local_slice(padsize=pad) if padded else F.real_local_slice()
    cs = F.complex_local_slice()
    dev = u.device
    loc = [torch.from_numpy(np.ascontiguousarray(v[sl])).to(dev).to(u.dtype) for v, sl in zip(vec, rs)]
    if dims == 3:
        torch.mul(loc[0][:, None, None] * loc[1][None, :, None], loc[2][None, None, :], out=u)
    else:
        torch.mul(loc[0][:, None], loc[1][None, :], out=u)
    fwd(u, fu, dealias)
    spec = [torch.from_numpy(np.ascontiguousarray(f[cs[ax]])).to(dev)
            for ax, f in enumerate(separable_spectra(vec, N, dealias, u.dtype == torch.float32))]
    num = torch.zeros((), dtype=torch.float64, device=dev)
    den = torch.zeros((), dtype=torch.float64, device=dev)
    rows = max(1, (256 << 20) // max(1, fu[0].numel() * 16))
    for i0 in range(0, fu.shape[0], rows):
        i1 = min(fu.shape[0], i0 + rows)
        if dims == 3:
            ref = spec[0][i0:i1, None, None] * spec[1][None, :, None] * spec[2][None, None, :]
        else:
            ref = spec[0][i0:i1, None] * spec[1][None, :]
        d = fu[i0:i1].to(torch.complex128) - ref
        num += (d.real ** 2 + d.imag ** 2).sum()
        den += (ref.real ** 2 + ref.imag ** 2).sum()
    return float(torch.sqrt(num / den).item())


def other_workload_names(main, P):
    """The BASELINE.json configurations besides the headline one that this rank count can run:
    config 4 (slab 1024^3 double, 3/2-rule), config 2 (slab 256^3 single), config 5a (line 16384^2 single), and from four
    ranks on the pencil ones -- 1024^3 double 'X', config 5b (pencil 'Y' 2048^3 single), config 3 (pencil 'X' 512^3 double
    on a 2 x 4 grid: eight ranks)."""
    names = ["slab1024_f64_32", "slab256_f32", "line16384_f32"]
    if P >= 4:
        names += ["pencilX1024_f64", "pencilY2048_f32"]
    if P == 8:
        names += ["pencilX512_f64"]
    return [n for n in names if n != main]


def quick_measure(m, comm, name, steps, torch, dist, P, rank):
    """Short device-resident measurement of one more workload (the loop of scripts/ab_multi.py, which produced the
    multi-GPU tables of DESIGN.md): forward parity against the closed form, 3 warm-up round trips, `steps` timed ones
    (CUDA events on the launching stream, L2 flushed between steps where the arrays would fit it, max over ranks),
    then one timed transform each way for the per-phase sums."""
    kind, N, prec, dealias, kw = WORKLOADS[name]
    F = make_transform(m, comm, name)
    fwd, inv = (F.fft2, F.ifft2) if kind == "line" else (F.fftn, F.ifftn)
    rshape = tuple(int(s) for s in (F.real_shape_padded() if dealias == "3/2-rule" else F.real_shape()))
    cshape = tuple(int(s) for s in F.complex_shape())
    rdt = torch.float64 if prec == "double" else torch.float32
    cdt = torch.complex128 if prec == "double" else torch.complex64
    g = torch.Generator(device="cuda").manual_seed(4321 + rank)
    u = torch.empty(rshape, dtype=rdt, device="cuda")
    fu = torch.empty(cshape, dtype=cdt, device="cuda")
    u2 = torch.empty_like(u)

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if P > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if P > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fe = None
    if not (kind == "line" and dealias):
        fe = reduce_max(forward_parity(F, kind, N, dealias, fwd, u, fu, torch))
    u.copy_(torch.rand(rshape, dtype=rdt, device="cuda", generator=g))
    for _ in range(3):
        fwd(u, fu, dealias)
        inv(fu, u2, dealias)
    # (a 3/2-rule round trip of a random padded field is not the identity -- the truncation drops its upper modes --
    # so only the plain transforms report one; forward parity covers both)
    rt = None if dealias else reduce_max(float((torch.linalg.vector_norm(u2 - u) / torch.linalg.vector_norm(u)).item()))
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda") if np.prod(N) * 4 <= 1 << 30 else None
    st = torch.cuda.current_stream()
    barrier()
    total = 0.0
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            fwd(u, fu, dealias)
            inv(fu, u2, dealias)
        e1.record(st)
        barrier()
        total = e0.elapsed_time(e1)
    else:
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            fwd(u, fu, dealias)
            inv(fu, u2, dealias)
            e1.record(st)
            barrier()
            total += e0.elapsed_time(e1)
    ms = reduce_max(total / steps)
    F.set_timing(True)
    fft_ms = ex_ms = ex_bytes = 0.0
    for call, a, b in ((fwd, u, fu), (inv, fu, u2)):
        call(a, b, dealias)
        torch.cuda.synchronize()
        for sname, sms, sbytes, _, _ in F.last_steps():
            if sname == "exchange":
                ex_ms += sms
                ex_bytes += sbytes
            else:
                fft_ms += sms
    F.set_timing(False)
    gshape = tuple(int(1.5 * n) if dealias == "3/2-rule" else n for n in N)
    cfg = describe(name, P)
    out = {"name": name, "workload": cfg["workload"], "decomposition": cfg["decomposition"], "l2": cfg["l2"], "steps": steps, "warmup": 3,
           "ms_per_step": round(ms, 4), "value": round(flops_roundtrip(gshape) / (ms * 1e-3) / 1e9, 1), "unit": UNIT,
           "dtype": "f64" if prec == "double" else "f32", "forward_rel_l2": fe, "roundtrip_rel_l2": rt,
           "sum_fft_ms": round(fft_ms, 4), "sum_exchange_ms": round(ex_ms, 4),
           "exchange_GBps_per_direction": round(ex_bytes / ex_ms / 1e6, 1) if ex_ms > 0 else None,
           "transport": getattr(F, "transport_used", None) if P > 1 else None, "workspace_bytes": F.workspace_bytes()}
    barrier()
    return out


def golden_check(m, comm, reduce_max):
    """Every golden file of the UNMODIFIED reference (tests/golden/*.npz, written by oracle/refshim/make_golden.py in the
    build container) whose rank count is this run's, through the classes with numpy arrays: ``fftn``, ``ifftn``, the
    3/2-rule both ways and the 2/3-rule, each rank comparing its block -- rel. L2, max over checks, files and ranks
    (``reduce_max``).  Tiny meshes: a few milliseconds per file; on N GPUs this is parity against the reference's own
    outputs over the real exchange.  No data-dependent control flow: every rank makes the same calls."""
    import glob
    P, r = comm.Get_size(), comm.Get_rank()
    L3 = np.array([2 * np.pi] * 3)
    worst = {"double": 0.0, "single": 0.0, "double_forward_3_2": 0.0, "single_forward_3_2": 0.0}
    names = []

    def err(got, ref):
        return float(np.linalg.norm((got - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300))

    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        z = np.load(path)
        meta = json.loads(str(z["meta"]))
        if meta["P"] != P:
            continue
        prec, N = meta["precision"], meta["N"]
        if meta["kind"] == "slab":
            F = m.Slab_R2C(np.array(N), L3, comm, prec, communication=meta["communication"])
            fwd, inv = F.fftn, F.ifftn
        elif meta["kind"] == "pencil":
            F = m.Pencil_R2C(np.array(N), L3, comm, prec, P1=meta["P1"], communication=meta["communication"],
                             alignment=meta["alignment"])
            fwd, inv = F.fftn, F.ifftn
        else:
            F = m.Line_R2C(np.array(N), L3[:2], comm, prec)
            fwd, inv = F.fft2, F.ifft2
        info = meta["ranks"][r]
        rs = tuple(slice(*x) for x in info["real_local_slice"])
        rps = tuple(slice(*x) for x in info["real_local_slice_padded"])
        cs = tuple(slice(*x) for x in info["complex_local_slice"])
        A, Cg, Ap = z["A"], z["C"], z["Ap"]
        e = [err(fwd(np.ascontiguousarray(A[rs]), np.zeros(Cg[cs].shape, dtype=Cg.dtype)), Cg[cs]),
             err(inv(np.ascontiguousarray(Cg[cs]), np.zeros(A[rs].shape, dtype=A.dtype)), z["A2"][rs])]
        Cin = Cg.copy()
        if meta["kind"] == "line":
            Cin[-N[0] // 2] = 0  # (tests/test_FFT.py:128: the line goldens' padded pair starts from this spectrum)
        e.append(err(inv(np.ascontiguousarray(Cin[cs]), np.zeros(Ap[rps].shape, dtype=Ap.dtype), dealias="3/2-rule"), Ap[rps]))
        if meta["has23"]:
            e.append(err(inv(np.ascontiguousarray(Cg[cs]), np.zeros(A[rs].shape, dtype=A.dtype), dealias="2/3-rule"), z["A23"][rs]))
        worst[prec] = max([worst[prec]] + e)
        e32 = err(fwd(np.ascontiguousarray(Ap[rps]), np.zeros(Cg[cs].shape, dtype=Cg.dtype), dealias="3/2-rule"), z["Cp"][cs])
        worst[prec + "_forward_3_2"] = max(worst[prec + "_forward_3_2"], e32)
        names.append(os.path.basename(path)[:-4])
        F = fwd = inv = None
    out = {k: reduce_max(v) for k, v in sorted(worst.items())}
    return {"files": names, "max_rel_l2": out,
            "what": "goldens of the unmodified reference with this rank count: fftn, ifftn, 3/2-rule inverse, 2/3-rule inverse "
                    "(max_rel_l2.double / .single) and the truncating 3/2-rule forward (…_forward_3_2), max over ranks"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import mpifft4py_b200 as m
    from mpifft4py_b200.comm import SelfComm, world

    P = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    if P > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = world()
    else:
        comm = SelfComm()
    name = args.workload
    kind, N, prec, dealias, kw = WORKLOADS[name]
    F = make_transform(m, comm, name)
    tuned = None
    if args.tune:  # let the planner pick among the opt-in kernels / schedules first (reported in config.tuning)
        tuned = m.tune.autotune(F, dealias=dealias, candidates=m.tune.CANDIDATES[args.tune])
        torch.cuda.empty_cache()
    fwd, inv = (F.fft2, F.ifft2) if kind == "line" else (F.fftn, F.ifftn)
    rshape = tuple(int(s) for s in (F.real_shape_padded() if dealias == "3/2-rule" else F.real_shape()))
    cshape = tuple(int(s) for s in F.complex_shape())
    rdt = torch.float64 if prec == "double" else torch.float32
    cdt = torch.complex128 if prec == "double" else torch.complex64
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    u = torch.rand(rshape, dtype=rdt, device="cuda", generator=g)
    fu = torch.empty(cshape, dtype=cdt, device="cuda")
    u2 = torch.empty_like(u)
    gshape = tuple(int(1.5 * n) if dealias == "3/2-rule" else n for n in N)
    flops = flops_roundtrip(gshape)
    small = np.prod(N) * 4 <= 1 << 30
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda") if small else None

    def barrier():
        if P > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # forward parity at this rank count (max over ranks), before the timed part; u is refilled afterwards
    fwd_err = None
    if dealias in (None, "3/2-rule") and not (kind == "line" and dealias):
        fe = torch.tensor([forward_parity(F, kind, N, dealias, fwd, u, fu, torch)], dtype=torch.float64, device="cuda")
        if P > 1:
            dist.all_reduce(fe, op=dist.ReduceOp.MAX)
        fwd_err = float(fe.item())
        u.copy_(torch.rand(rshape, dtype=rdt, device="cuda", generator=g))

    def step():
        fwd(u, fu, dealias)
        inv(fu, u2, dealias)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    st = torch.cuda.current_stream()
    times = []
    barrier()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(args.steps):
            step()
        e1.record(st)
        barrier()
        total_ms = e0.elapsed_time(e1)
    else:
        total_ms = 0.0
        for _ in range(args.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            step()
            e1.record(st)
            barrier()
            total_ms += e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if P > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tms.item()) / args.steps
    value = flops / (ms_per_step * 1e-3) / 1e9
    k1, x1 = F.last_launches()
    err = (torch.linalg.vector_norm(u2 - u) / torch.linalg.vector_norm(u)).item()

    # ---- roofline: per-pass device times from events on the launching stream --------------------
    F.set_timing(True)
    acc = {}
    reps = 3

    def record(direction, rep):
        # the chunks of a pipelined pass share a pass id: add their times and bytes
        for s in F.last_steps():
            a = acc.setdefault((direction, s[4], s[0], s[3]), [0.0, 0.0])
            a[0] += s[1]
            if rep == 0:
                a[1] += s[2]

    # Each timed transform is the last of a run of back-to-back round trips without host synchronisation, so it is
    # measured at the clocks the timed K-step region saw (a transform timed alone after a synchronize runs at boost
    # clocks: round 1's per-pass sum was 6 % under its step time for exactly that reason -- sw_power_cap, BENCH_r01).
    lead = 0 if flush is not None else 3
    for rep in range(reps):
        for _ in range(lead):
            step()
        if flush is not None:
            flush.fill_(1)
        fwd(u, fu, dealias)
        torch.cuda.synchronize()
        record("fwd", rep)
        for _ in range(lead):
            step()
        if flush is not None:
            flush.fill_(1)
        inv(fu, u2, dealias)
        torch.cuda.synchronize()
        record("inv", rep)
    F.set_timing(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    passes = []
    for (d, i, ty, ln), (ms, by) in sorted(acc.items()):
        ms /= reps
        passes.append({"dir": d, "step": i, "type": ty, "len": ln, "ms": round(ms, 4), "bytes": by,
                       "GBps": round(by / (ms * 1e-3) / 1e9, 1) if ms > 0 else None})
    ffts = [p for p in passes if p["type"] != "exchange"]
    dom = max(ffts, key=lambda p: p["ms"])
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(name, {}).get("%s_%d_%s_%d" % (dom["dir"], dom["step"], dom["type"], dom["len"])) if P == 1 else None
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": "%s %s n=%d (step %d)" % (dom["dir"], dom["type"], dom["len"], dom["step"]),
                "achieved": dom["GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": round(dom["GBps"] / hbm_peak, 4),
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": dom["bytes"], "ms": dom["ms"],
                "passes": passes,
                "pass_timing": "CUDA events around each pass of a transform that ends a run of %d back-to-back round trips" % lead,
                "sum_fft_ms": round(sum(p["ms"] for p in ffts), 4),
                "sum_exchange_ms": round(sum(p["ms"] for p in passes if p["type"] == "exchange"), 4)}
    # whole-step roofline: time the algorithmic bytes need at the HBM figure plus the
    # exchanged bytes at the link figure, serial and fully overlapped, against the measured step
    t_hbm = sum(p["bytes"] for p in ffts) / (hbm_peak * 1e9) * 1e3
    t_link = sum(p["bytes"] for p in passes if p["type"] == "exchange") / 770e9 * 1e3
    roofline["step_model"] = {"hbm_ms": round(t_hbm, 4), "link_ms": round(t_link, 4), "serial_ms": round(t_hbm + t_link, 4),
                              "overlapped_ms": round(max(t_hbm, t_link), 4),
                              "frac_of_serial": round((t_hbm + t_link) / ms_per_step, 4),
                              "frac_of_overlapped": round(max(t_hbm, t_link) / ms_per_step, 4),
                              "link_peak_GBps_per_direction": 770.0}
    xs = [p for p in passes if p["type"] == "exchange" and p["ms"] > 0]
    if xs:
        ach = sum(p["bytes"] for p in xs) / sum(p["ms"] for p in xs) / 1e6
        roofline["nvlink"] = {"achieved_GBps_per_direction": round(ach, 1), "peak": 770.0, "unit": "GB/s",
                              "frac": round(ach / 770.0, 3), "frac_of_nominal_900": round(ach / 900.0, 3),
                              "peak_source": "measured peer copy 770 GB/s per direction (B200_PROFILING.md); nominal NVLink 5: 900",
                              "note": "exchange steps are timed on the communication stream while FFT passes of the "
                                      "neighbouring chunks run beside them (they share HBM)"}

    # ---- e2e: numpy API, pinned host buffers, H2D + D2H inside the timed region -------------------
    e2e = None
    if not args.no_e2e:
        try:
            del u2
            torch.cuda.empty_cache()
            rnp, cnp = (np.float64, np.complex128) if prec == "double" else (np.float32, np.complex64)
            hu = m.empty(rshape, dtype=rnp)
            hf = m.empty(cshape, dtype=cnp)
            hu[...] = np.random.default_rng(1234 + rank).random(rshape[-1], dtype=np.float64).astype(rnp)  # broadcast rows
            nst = max(1, args.e2e_steps or args.steps)
            fwd(hu, hf, dealias)
            inv(hf, hu, dealias)  # warm-up (allocates the staging buffers)
            barrier()
            t0 = time.perf_counter()
            for _ in range(nst):
                fwd(hu, hf, dealias)   # H2D(u) + transform + D2H(fu)
                inv(hf, hu, dealias)   # H2D(fu) + transform + D2H(u)
            barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / nst], dtype=torch.float64, device="cuda")
            if P > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            nb = hu.nbytes + hf.nbytes
            e2e = {"value": flops / float(dt.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(nb),
                   "d2h_bytes_per_step": int(nb), "ms_per_step": float(dt.item()) * 1e3, "steps": nst,
                   "api": "numpy arrays in pinned host memory through %s.fftn/ifftn" % type(F).__name__}
        except Exception as e:  # noqa: BLE001
            e2e = {"value": None, "unit": UNIT, "error": repr(e)[:200]}

    cpu = None
    if rank == 0 and P == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(name)
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "error": repr(e)[:200]}

    if P > 1:
        dist.barrier()
    cfg = describe(name, P)
    # plan options set through the environment: a line measured with any of them says so
    tuning = {k: os.environ[k] for k in ("B200FFT_LAYOUT", "B200FFT_TRANSPORT", "B200FFT_PIPELINE", "B200FFT_CHUNKS",
                                          "B200FFT_FLAG_DMA") if os.environ.get(k)}
    if tuned is not None:
        tuning["planner"] = {"effort": args.tune, "chosen": tuned["chosen"],
                             "candidates": [{k: c.get(k) for k in ("name", "seconds", "ok")} for c in tuned["candidates"]]}
    if tuning:
        cfg["tuning"] = tuning
    if P > 1:
        ex = [s for s in F.last_steps() if s[0] == "exchange"]
        cfg["exchange"] = {"transport": getattr(F, "transport_used", "nccl"),
                           "pipelined_chunks": len(ex) // max(1, len({s[4] for s in ex}))}
    workspace = F.workspace_bytes()

    # ---- the other BASELINE configurations this rank count can run: short device-resident lines with forward parity ----
    others = None
    if not args.no_others:
        # everything of the headline workload goes first: its arrays, staging buffers, pinned host arrays and plan
        hu = hf = u = fu = u2 = flush = fwd = inv = F = None   # (fwd / inv are bound methods: they hold the object)
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        others, t_start = [], time.perf_counter()
        for oname in other_workload_names(name, P):
            spent = torch.tensor([time.perf_counter() - t_start], dtype=torch.float64, device="cuda")
            if P > 1:
                dist.all_reduce(spent, op=dist.ReduceOp.MAX)  # every rank takes the same decision
            if float(spent.item()) > args.others_budget:
                others.append({"name": oname, "skipped": "time budget of the extra workloads used up"})
                continue
            try:
                others.append(quick_measure(m, comm, oname, args.others_steps, torch, dist, P, rank))
            except Exception as e:  # noqa: BLE001 - the headline line must not depend on the extras
                others.append({"name": oname, "error": repr(e)[:300]})
            gc.collect()
            torch.cuda.empty_cache()

    # ---- parity against the reference's own outputs at THIS rank count (stored goldens; milliseconds) --------------------
    golden = None
    if not args.no_golden:
        def rmax(x):
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            if P > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        try:
            golden = golden_check(m, comm, rmax)
        except Exception as e:  # noqa: BLE001
            golden = {"error": repr(e)[:300]}
        import gc
        gc.collect()

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": P, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64" if prec == "double" else "f32", "data": "synthetic", "config": cfg,
               "roundtrip_rel_l2": err, "forward_rel_l2": fwd_err,
               "forward_parity": "separable field a(x)b(y)c(z) vs the outer product of numpy 1D FFTs cut by complex_local_slice(), max over ranks",
               "gpu_launches": int(k1) * 2 * args.steps if kind else 0,
               "kernels_per_transform": int(k1), "nccl_groups_per_transform": int(x1),
               "roofline": roofline, "clocks": clocks, "e2e": e2e, "cpu_baseline": cpu,
               "workspace_bytes": workspace, "other_workloads": others, "reference_goldens": golden}
        print(json.dumps(out))
    if P > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="slab1024_f64", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the end-to-end leg (0 = --steps)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true",
                    help="skip the short lines of the other BASELINE configurations (key other_workloads)")
    ap.add_argument("--no-golden", action="store_true", help="skip the check against the stored outputs of the unmodified reference")
    ap.add_argument("--others-steps", type=int, default=5)
    ap.add_argument("--others-budget", type=float, default=60.0, help="seconds after which no further extra workload is started")
    ap.add_argument("--tune", default=None, choices=["measure", "patient"],
                    help="run mpifft4py_b200.tune.autotune on the workload first (default: the library's defaults)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world_size:
        if world_size == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
