#!/usr/bin/env python
"""A/B of exchange transports / pipelines on N GPUs in ONE process group (one torch import, one NCCL bootstrap):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/ab_multi.py [--steps 10] [--workloads slab1024_f64,...] [--configs name,name]

For every (workload, configuration): a fresh transform object with the configuration's plan attributes,
forward parity against the closed form of a separable field (bench.forward_parity: catches permutation mistakes a
round trip cannot), round-trip error, K timed round trips (CUDA events, max over ranks) and the per-pass table.
One JSON line per case on stdout (rank 0), a table on stderr.  Numbers that are reported come from bench.py.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mpifft4py_b200 as m  # noqa: E402
from mpifft4py_b200.comm import world  # noqa: E402

# name -> plan attributes (transport, exchange_pipeline, exchange_chunks)
CONFIGS = {
    "default": {},
    "nccl_c1": {"transport": "nccl", "exchange_chunks": 1},
    "nccl_c2": {"transport": "nccl", "exchange_chunks": 2},
    "nccl_c4": {"transport": "nccl", "exchange_chunks": 4},
    "p2p_c1": {"transport": "p2p", "exchange_chunks": 1},
    "p2p_c2": {"transport": "p2p", "exchange_chunks": 2},
    "p2p_c4": {"transport": "p2p", "exchange_chunks": 4},
    "p2p_c8": {"transport": "p2p", "exchange_chunks": 8},
    "store_c1": {"transport": "store", "exchange_chunks": 1},
    "store_c2": {"transport": "store", "exchange_chunks": 2},
    "store_c4": {"transport": "store", "exchange_chunks": 4},
    "nccl_kz2": {"transport": "nccl", "exchange_pipeline": "kz", "exchange_chunks": 2},
    "nccl_kz4": {"transport": "nccl", "exchange_pipeline": "kz", "exchange_chunks": 4},
    "p2p_kz2": {"transport": "p2p", "exchange_pipeline": "kz", "exchange_chunks": 2},
    "p2p_kz4": {"transport": "p2p", "exchange_pipeline": "kz", "exchange_chunks": 4},
    "store_kz2": {"transport": "store", "exchange_pipeline": "kz", "exchange_chunks": 2},
    "store_kz4": {"transport": "store", "exchange_pipeline": "kz", "exchange_chunks": 4},
    "store_kz8": {"transport": "store", "exchange_pipeline": "kz", "exchange_chunks": 8},
}
SLAB_SET = ["default", "nccl_c1", "nccl_c2", "p2p_c1", "p2p_c2", "p2p_c4", "p2p_c8", "store_c1", "store_c2", "store_c4",
            "nccl_kz2", "p2p_kz2", "p2p_kz4", "store_kz2", "store_kz4", "store_kz8"]
OTHER_SET = ["default", "nccl_c1", "nccl_c2", "nccl_c4", "p2p_c1", "p2p_c2", "p2p_c4", "store_c1", "store_c2"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--workloads", default="slab1024_f64")
    ap.add_argument("--configs", default="")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = world()
    P, rank = comm.Get_size(), comm.Get_rank()
    for name in args.workloads.split(","):
        kind, N, prec, dealias, kw = bench.WORKLOADS[name]
        names = args.configs.split(",") if args.configs else (SLAB_SET if kind == "slab" else OTHER_SET)
        try:
            F0 = bench.make_transform(m, comm, name)
        except Exception as e:  # noqa: BLE001 - e.g. a pencil grid this rank count does not allow
            if rank == 0:
                sys.stderr.write("%-18s skipped: %r\n" % (name, e))
            continue
        rshape = tuple(int(s) for s in (F0.real_shape_padded() if dealias == "3/2-rule" else F0.real_shape()))
        cshape = tuple(int(s) for s in F0.complex_shape())
        del F0
        rdt = torch.float64 if prec == "double" else torch.float32
        cdt = torch.complex128 if prec == "double" else torch.complex64
        g = torch.Generator(device="cuda").manual_seed(1234 + rank)
        u = torch.empty(rshape, dtype=rdt, device="cuda")
        fu = torch.empty(cshape, dtype=cdt, device="cuda")
        u2 = torch.empty_like(u)
        flops = bench.flops_roundtrip(tuple(int(1.5 * n) if dealias == "3/2-rule" else n for n in N))
        for cname in names:
            attrs = CONFIGS[cname]
            out = {"workload": name, "config": cname, "attrs": attrs, "n_gpus": P}
            try:
                F = bench.make_transform(m, comm, name)
                for k, v in attrs.items():
                    setattr(F, k, v)
                fwd, inv = (F.fft2, F.ifft2) if kind == "line" else (F.fftn, F.ifftn)
                fe = torch.tensor([bench.forward_parity(F, kind, N, dealias, fwd, u, fu, torch)], dtype=torch.float64, device="cuda")
                dist.all_reduce(fe, op=dist.ReduceOp.MAX)
                u.copy_(torch.rand(rshape, dtype=rdt, device="cuda", generator=g))
                for _ in range(3):
                    fwd(u, fu, dealias)
                    inv(fu, u2, dealias)
                rt = torch.linalg.vector_norm(u2 - u) / torch.linalg.vector_norm(u)
                rt = rt.to(torch.float64).reshape(1)
                dist.all_reduce(rt, op=dist.ReduceOp.MAX)
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    fwd(u, fu, dealias)
                    inv(fu, u2, dealias)
                e1.record()
                dist.barrier()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
                F.set_timing(True)
                fft_ms = ex_ms = ex_bytes = 0.0
                for call, a, b2 in ((fwd, u, fu), (inv, fu, u2)):
                    call(a, b2, dealias)
                    torch.cuda.synchronize()
                    for s in F.last_steps():
                        if s[0] == "exchange":
                            ex_ms += s[1]
                            ex_bytes += s[2]
                        else:
                            fft_ms += s[1]
                F.set_timing(False)
                out.update({"ms_per_round_trip": round(ms, 3), "GFLOPs": round(flops / ms / 1e6), "forward_rel_l2": float(fe.item()),
                            "roundtrip_rel_l2": float(rt.item()), "transport_used": getattr(F, "transport_used", None),
                            "sum_fft_ms": round(fft_ms, 3), "sum_exchange_ms": round(ex_ms, 3),
                            "exchange_GBps_per_direction": round(ex_bytes / ex_ms / 1e6, 1) if ex_ms > 0 else None,
                            "workspace_bytes": F.workspace_bytes()})
                del F
            except Exception as e:  # noqa: BLE001
                out["error"] = repr(e)[:300]
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.empty_cache()
            if rank == 0:
                print(json.dumps(out), flush=True)
                if "error" in out:
                    sys.stderr.write("%-18s %-12s ERROR %s\n" % (name, cname, out["error"]))
                else:
                    sys.stderr.write("%-18s %-12s %8.3f ms  fwd %.1e  rt %.1e  fft %.2f  exch %.2f ms (%s GB/s/dir)  [%s]\n" % (
                        name, cname, out["ms_per_round_trip"], out["forward_rel_l2"], out["roundtrip_rel_l2"], out["sum_fft_ms"],
                        out["sum_exchange_ms"], out["exchange_GBps_per_direction"], out["transport_used"]))
        del u, fu, u2
        torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
