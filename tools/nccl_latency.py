#!/usr/bin/env python3
"""Small-message collective latency between the GPUs of one box (evidence for DESIGN.md: a per-merge exchange
would cost more than a whole resident merge step). torchrun --nproc-per-node N tools/nccl_latency.py"""
import os, json
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
res = {}
for name, n in (("allreduce_64B", 16), ("allreduce_4KB", 1024), ("allgather_4KB", 1024)):
    x = torch.ones(n, dtype=torch.int32, device="cuda")
    outs = torch.empty(n * world, dtype=torch.int32, device="cuda")
    fn = (lambda: dist.all_reduce(x)) if name.startswith("allreduce") else (lambda: dist.all_gather_into_tensor(outs, x))
    for _ in range(50): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 2000
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name + "_us"] = round(float(t.item()), 2)
if rank == 0:
    print(json.dumps({"world": world, "back_to_back_stream_latency": res}))
dist.destroy_process_group()
