"""torchrun --nproc-per-node N tools/multi_gpu_check.py : the fused GEMM->peer gather equals the NCCL all-gather."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder
from zoomearth_b200.sharding import PeerGather
from zoomearth_b200.synthetic import random_vision_state_dict

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sd = random_vision_state_dict(0, device=dev, depth=2)
fv = FusedVisual(sd, device=dev, dtype=torch.float16, depth=2, fullatt=[1])
enc = ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=200704, device=dev))
g = torch.Generator(device=dev).manual_seed(100 + rank)
imgs = [torch.randint(0, 256, (700, 900, 3), generator=g, dtype=torch.uint8, device=dev) for _ in range(3)]
emb, grid, _ = enc.encode(imgs, None)
T = emb.shape[0]
ref = torch.empty((world * T, emb.shape[1]), dtype=emb.dtype, device=dev)
dist.all_gather_into_tensor(ref, emb)
pg = PeerGather(world * T, emb.shape[1], torch.float16, dev)
pg.buffer.zero_()
pg.barrier()
enc.encode(imgs, None, gather=pg, gather_row=rank * T)
pg.barrier()
torch.cuda.synchronize()
same = torch.equal(pg.buffer, ref)
print(f"rank {rank}: fused gather == nccl all_gather: {same}  (T={T}, peers={len(pg.peer_ptrs)})", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if same else 1)
