"""torchrun --nproc-per-node N tools/multi_gpu_check.py : correctness of the fused multi-GPU data paths against NCCL.

  1. uniform shards: the merger GEMM's peer stores (zv_visual_forward_gather) == all_gather_into_tensor, bitwise;
  2. the same with two gather buffers and the barrier deferred to a side stream (PeerGather(double_buffer=True)), 3 steps;
  3. ragged shards: LPT-partitioned mixed-size crops through zv_visual_forward_gather_rows == the NCCL ragged gather
     (sharding.gather_embeddings: counts + padded all-gather + permutation), bitwise, and identical on every rank.
Prints one line per check per rank; exit code 0 only if every check holds.  Logs are committed under profiles/.
"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder, sharding, synthetic
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
ok = True

# 1. uniform shards, single buffer
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
ok &= same
print(f"rank {rank}: fused gather == nccl all_gather: {same}  (T={T}, peers={len(pg.peer_ptrs)})", flush=True)

# 2. double-buffered, deferred barrier: three steps with different inputs
pg2 = PeerGather(world * T, emb.shape[1], torch.float16, dev, double_buffer=True)
refs = []
for step in range(3):
    gs = torch.Generator(device=dev).manual_seed(1000 * step + rank)
    im = [torch.randint(0, 256, (700, 900, 3), generator=gs, dtype=torch.uint8, device=dev) for _ in range(3)]
    e, _, _ = enc.encode(im, None)
    r = torch.empty((world * T, e.shape[1]), dtype=e.dtype, device=dev)
    dist.all_gather_into_tensor(r, e)
    refs.append(r)
    ready = pg2.begin_step()
    if ready is not None:
        s2 = torch.equal(ready, refs[step - 1])
        ok &= s2
        print(f"rank {rank}: double-buffered gather, step {step - 1} ready at step {step}: {s2}", flush=True)
    enc.encode(im, None, gather=pg2, gather_row=rank * T)
    pg2.end_step()
s2 = torch.equal(pg2.finish(), refs[-1])
ok &= s2
print(f"rank {rank}: double-buffered gather, last step after finish(): {s2}", flush=True)

# 3. ragged shards: mixed-size crops, LPT partition, per-row peer stores
enc3 = ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=1003520, device=dev))
gp = torch.Generator(device=dev).manual_seed(7)
pool = [torch.randint(0, 256, (1600, 1600, 3), generator=gp, dtype=torch.uint8, device=dev) for _ in range(2)]
boxes, index = synthetic.mixed_crop_boxes(23, n_images=2, img=1600, lo=200, hi=1200, seed=3)
from zoomearth_b200 import geometry
_, _, g3 = geometry.geometry(enc3.processor._cfg(), np.array([[1600, 1600]] * 23, np.int32), boxes.astype(np.float64))
tokens = (g3[:, 1] * g3[:, 2]) // 4
pg3 = PeerGather(int(tokens.sum()), 2048, torch.float16, dev)
pg3.buffer.zero_()
pg3.barrier()
out, tok, parts = sharding.encode_sharded(enc3, pool, boxes, index, pg3, max_patches=6000)
pg3.barrier()
torch.cuda.synchronize()
mine = parts[rank]
if mine:
    # the same micro-batches without the gather, then NCCL's ragged gather
    embs = []
    lb, li = [boxes[i] for i in mine], [int(index[i]) for i in mine]
    for grp in enc3.micro_batches(pool, lb, li, 6000):
        used = sorted({li[i] for i in grp}); loc = {k: j for j, k in enumerate(used)}
        e, _, _ = enc3.encode([pool[k] for k in used], [lb[i] for i in grp], image_index=[loc[li[i]] for i in grp])
        embs.append(e)
    local_emb = torch.cat(embs)
else:
    local_emb = torch.zeros((0, 2048), dtype=torch.float16, device=dev)
ref3, counts = sharding.gather_embeddings(local_emb, [int(tokens[i]) for i in mine], parts)
s3 = torch.equal(out, ref3) and counts.tolist() == tokens.tolist()
ok &= s3
print(f"rank {rank}: ragged fused gather == nccl ragged gather: {s3}  ({len(mine)} of 23 crops here, {int(tokens.sum())} rows)", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
