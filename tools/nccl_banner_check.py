#!/usr/bin/env python
"""One-rank NCCL init + all_reduce: shows on which stream NCCL prints its banner (run under NCCL_DEBUG=VERSION)."""
import os
import torch
import torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.ones(4, device="cuda")
dist.all_reduce(t)
torch.cuda.synchronize()
print("ALLREDUCE_OK", t.tolist(), flush=True)
dist.destroy_process_group()
