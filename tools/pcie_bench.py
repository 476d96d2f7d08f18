"""Concurrent pinned host<->device copy ceiling of the box, one process per GPU (torchrun), with and without
binding each process to its GPU-local host cores. Sizes = the bench's e2e step (cartpole T=101, B=4096:
29.9 MB in, 98.7 MB out per GPU).   torchrun --nproc-per-node N tools/pcie_bench.py [--bind]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import bind_near_gpu  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
binding = bind_near_gpu(local) if "--bind" in sys.argv else None
if world > 1:
    dist.init_process_group("gloo")
H2D, D2H = 29_917_184, 98_697_216
hin = torch.empty(H2D, dtype=torch.uint8).pin_memory()
hout = torch.empty(D2H, dtype=torch.uint8).pin_memory()
hin.fill_(1)
hout.fill_(0)
din = torch.empty(H2D, dtype=torch.uint8, device="cuda")
dout = torch.ones(D2H, dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
for n_active in (1, 2, 4, 8):
    if n_active > world:
        break
    active = rank < n_active  # the other ranks only join the barriers
    res = {}
    for mode in ("d2h", "h2d", "both"):
        for it in range(2):  # warm-up pass, timed pass
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            if active:
                for _ in range(20):
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s_in):
                            din.copy_(hin, non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s_out):
                            hout.copy_(dout, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 20
        t = torch.tensor([dt if active else 0.0], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        byt = (H2D if mode != "d2h" else 0) + (D2H if mode != "h2d" else 0)
        res[mode] = {"ms_max_over_ranks": 1e3 * float(t), "GBs_per_gpu": byt / float(t) / 1e9, "GBs_aggregate": n_active * byt / float(t) / 1e9}
    if rank == 0:
        print(json.dumps({"n_gpus": n_active, "launched": world, "bind": "--bind" in sys.argv, "binding_rank0": binding,
                          "host_cpus": os.cpu_count(), **res}), flush=True)
if world > 1:
    dist.destroy_process_group()
