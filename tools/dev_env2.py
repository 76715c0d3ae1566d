"""Dev-only: closed-loop env-steps/s of this process (optionally after initialising NCCL), to debug multi-rank slowness."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if "RANK" in os.environ and os.environ.get("USE_DIST", "1") == "1":
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    t = torch.ones(1, device=f"cuda:{local}"); dist.all_reduce(t)
for rep in range(2):
    t0 = time.time()
    r, take = bench.env_steps_per_sec(local, 1, 8192, 20, 3 + local)
    print(f"rank {os.environ.get('RANK','-')} dev {local} rep {rep}: {r:.0f} env-steps/s take {take:.4f} wall {time.time()-t0:.1f}s", flush=True)
