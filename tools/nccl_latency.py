"""Latency of the small collectives the row-sharded search uses (torchrun, NCCL): device-timed, max over ranks.
    torchrun --nproc-per-node N tools/nccl_latency.py"""
import json
import os

import torch
import torch.distributed as dist


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    out = {}
    for name, nbytes in (("16KB", 16 << 10), ("1.6MB", 1638400), ("3.3MB", 3276800)):
        src = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        dst = torch.zeros(nbytes * world, dtype=torch.uint8, device=dev)
        for kind in ("all_gather", "all_reduce_min"):
            if kind == "all_reduce_min" and nbytes > (64 << 10):
                continue
            f32 = src.view(torch.float32)

            def op():
                if kind == "all_gather":
                    dist.all_gather_into_tensor(dst, src)
                else:
                    dist.all_reduce(f32, op=dist.ReduceOp.MIN)
            for _ in range(10):
                op()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 50
            e0.record()
            for _ in range(reps):
                op()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[f"{kind}_{name}_us"] = round(float(t.item()), 1)
    if dist.get_rank() == 0:
        print(json.dumps({"world": world, **out}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
