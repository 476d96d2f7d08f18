"""BASELINE config 5: cartpole T in {51, 201, 1001} x batch 1K .. 256K sharded over the GPUs of one box, weak and
strong scaling, device-timed fused Jacobian+Hessian pass (max over ranks), one JSON line per case with the CPU
baseline (oracle C twin, all host cores, bounded sample) beside it.

    python tools/config5_sweep.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/config5_sweep.py

weak:   B problems PER GPU (B in 1K, 4K, 16K, 64K, 256K where they fit the per-GPU memory cap)
strong: B problems IN TOTAL, contiguous shards of ceil(B/N) (sharding.partition)
Inputs are drawn on the device with the SURVEY 8(d) distributions (angles U(-pi, pi), velocities N(0,1),
u ~ U(-3, 3), lambda ~ N(0,1), sigma = 1, w = [x1; xT] + N(0, 0.1^2)): a 256K x T=1001 batch is 82 GB, too much
to stage through host memory. One launch of such a case streams far more than the 126 MB L2, so no rotation is needed;
small cases rotate 3 buffer sets."""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dto_b200 as D  # noqa: E402
from dto_b200 import sharding  # noqa: E402
from dto_b200.evaluator import A_LAMBDA, A_SIGMA, A_W, A_Z, K_JAC_HESS  # noqa: E402
from dto_b200.sqp import _CudaArray  # noqa: E402
from examples import models as M  # noqa: E402

MEM_CAP = float(os.environ.get("DTO_SWEEP_MEM_GB", "120")) * 1e9
TS = [int(t) for t in os.environ.get("DTO_SWEEP_T", "51,201,1001").split(",")]
BS = [int(b) for b in os.environ.get("DTO_SWEEP_B", "1024,4096,16384,65536,262144").split(",")]


def fill_on_device(nlp, T, B, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    view = lambda arr, shape: torch.as_tensor(_CudaArray(nlp.device_pointer(arr, 0), shape), device=dev)  # noqa: E731
    N_z = nlp.num_variables
    z = view(A_Z, (B, N_z))
    body = z[:, :(T - 1) * 5].view(B, T - 1, 5)
    body[:, :, 0:2] = (torch.rand((B, T - 1, 2), generator=g, device=dev, dtype=torch.float64) * 2 - 1) * math.pi
    body[:, :, 2:4] = torch.randn((B, T - 1, 2), generator=g, device=dev, dtype=torch.float64)
    body[:, :, 4] = (torch.rand((B, T - 1), generator=g, device=dev, dtype=torch.float64) * 2 - 1) * 3.0
    z[:, -4:-2] = (torch.rand((B, 2), generator=g, device=dev, dtype=torch.float64) * 2 - 1) * math.pi
    z[:, -2:] = torch.randn((B, 2), generator=g, device=dev, dtype=torch.float64)
    view(A_LAMBDA, (B, nlp.num_constraint)).normal_(generator=g)
    view(A_SIGMA, (B,)).fill_(1.0)
    w = view(A_W, (B, 8))
    w.normal_(0.0, 0.1, generator=g)
    w[:, 5] += math.pi


def cpu_baseline(T, threads):
    from bench import build_c_baseline, time_cpu
    from util import make_inputs
    co, mo = build_c_baseline(T)
    Bs = max(threads * 4, min(512, int(2.0e5 / T)))
    z, lam, sigma, w = make_inputs("cartpole", mo, co.NZ, co.NC, co.NW, Bs, config=5)
    v, n, reps, dt = time_cpu(co, z, lam, sigma, w, threads, budget_s=3.0)
    return {"value": v, "unit": "knot-evals/s", "cores": threads, "kind": "port", "sample": f"{n} problems x T={T}, {reps} passes, {dt:.1f} s"}


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    threads = max(1, len(os.sched_getaffinity(0)))
    for T in TS:
        model = M.build_cartpole(D, T=T, evaluate_hessian=True, parameterized=True)
        base = cpu_baseline(T, threads) if rank == 0 else None
        for mode in ("weak", "strong"):
            for Bn in BS:
                B = Bn if mode == "weak" else sharding.partition(Bn, world)[rank][1]
                if mode == "strong" and world == 1:
                    continue  # identical to the weak row
                solver = D.solver_from(model, batch=max(B, 1), devices=[local])
                nlp = solver.nlp
                per_problem = nlp.algorithmic_bytes_per_problem()
                too_big = torch.tensor([1.0 if per_problem * B > MEM_CAP else 0.0], device=dev)
                if world > 1:
                    dist.all_reduce(too_big, op=dist.ReduceOp.MAX)
                if too_big.item() > 0 or B == 0:
                    nlp.close()
                    continue
                R = 3 if per_problem * B * 3 < 4e9 else 1
                nl = [nlp] + [nlp.new_batch() for _ in range(R - 1)]
                st = torch.cuda.current_stream(dev)
                for i, n_ in enumerate(nl):
                    n_.set_stream(st.cuda_stream)
                    fill_on_device(n_, T, B, 20261017 + 5000 + 17 * rank + i, dev)
                steps = int(max(3, min(100, 2e10 / (per_problem * B))))
                for i in range(3):
                    nl[i % R].launch(K_JAC_HESS)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for i in range(steps):
                    nl[i % R].launch(K_JAC_HESS)
                e1.record(st)
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                ms = float(ms.item())
                total = Bn * world if mode == "weak" else Bn
                if rank == 0:
                    gbs = per_problem * total / ms / 1e6
                    print(json.dumps({"config": 5, "model": "cartpole", "T": T, "scaling": mode, "n_gpus": world, "B_total": total,
                                      "B_per_gpu": B, "ms": ms, "evals_per_s": total * T / ms * 1e3, "GBs_aggregate": gbs,
                                      "frac_hbm_per_gpu": gbs / world / 6549.4, "steps": steps, "rotating_sets": R,
                                      "cpu_baseline": base, "speedup_vs_cpu": total * T / ms * 1e3 / base["value"]}), flush=True)
                for n_ in nl:
                    n_.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
