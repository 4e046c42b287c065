"""Development tool (torchrun): the bench's ESS sweep repeated on N GPUs; prints the walls of every repetition."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
import drjacoby_b200 as drj

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def barrier():
    torch.cuda.synchronize(); dist.barrier()
for rep in range(3):
    r = bench.ess_sweep(drj, local, rank, world, barrier)
    if rank == 0:
        print(rep, "walls", [round(w, 4) for w in r["wall_s_runs"]], "ESS/s %.4g" % r["value"], flush=True)
dist.barrier()
dist.destroy_process_group()
