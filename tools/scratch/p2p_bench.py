import os, sys, ctypes as C
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feltor_b200.dist import Comm
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = Comm.from_torch_distributed()
rec = torch.zeros(41, dtype=torch.int64, device="cuda")
rec[20] = 12345 + comm.rank
for _ in range(20):
    comm.allreduce_dot(rec)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    comm.allreduce_dot(rec)
e1.record(); torch.cuda.synchronize()
print("rank", comm.rank, "allreduce_dot %.2f us per call" % (e0.elapsed_time(e1) / 200 * 1e3), "NO_P2P=", os.environ.get("DGB_NO_P2P"), flush=True)
dist.destroy_process_group()
