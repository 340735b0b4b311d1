#!/usr/bin/env python
"""Throughput of the ragged BASELINE.json workloads (configs[2], [3], [4]) through the public
fast path (ZoomEncoder.encode_batched: cut_image rule -> K1 -> tower), one JSON line per config.

    python tools/bench_configs.py --config 3 4 5            # one GPU, bounded sizes by default (--full for SURVEY sizes)
    torchrun --nproc-per-node N tools/bench_configs.py --config 4     # crops sharded by LPT partition, NCCL ragged gather

These are parity-test workloads, not the bench line (bench.py times configs[1]); the numbers go under profiles/.
Source images stay resident (16 synthetic 5000x5000 uint8 images per GPU), crops are encoded in micro-batches of
at most --max-patches patches, timing is CUDA events around the whole pass after one warm-up pass, max over ranks.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KCLASS = {"k1_hpass": 0, "k1_vpass": 1, "gemm_store": 2, "gemm_qkv": 3, "gemm_resid": 4, "gemm_swiglu": 5,
          "gemm_gelu": 6, "gemm_scatter": 7, "attn_window": 8, "attn_full": 9, "rmsnorm": 10, "gather": 11}


def flops(grid):
    """Algorithmic FLOPs of the tower for these grids: linear, window attention, full attention."""
    g = np.asarray(grid, np.int64)
    S = g[:, 1] * g[:, 2]
    f_lin = float((S * 1262940160 + (S // 4) * 73400320).sum())
    f_full = float((4 * 4 * S.astype(np.float64) ** 2 * 1280).sum())
    f_win = 0.0
    for _, gh, gw in g:
        lh, lw = gh // 2, gw // 2
        ny = [4 * min(4, lh - y) for y in range(0, lh, 4)]
        nx = [min(4, lw - x) for x in range(0, lw, 4)]
        f_win += 28 * sum(4.0 * (a * b) ** 2 * 1280 for a in ny for b in nx)
    return f_lin, f_win, f_full


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, nargs="+", default=[3, 4, 5])
    ap.add_argument("--full", action="store_true", help="SURVEY sizes (256 questions / 1024 crops / 128 max-res crops)")
    ap.add_argument("--max-patches", type=int, default=400_000)
    ap.add_argument("--pool", type=int, default=16, help="resident source images per GPU")
    args = ap.parse_args()

    import torch.distributed as dist
    from zoomearth_b200 import FusedImageProcessor, FusedVisual, ZoomEncoder, _lib, sharding, synthetic

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stdout_fd = None
    if world > 1:
        sys.stdout.flush()
        stdout_fd = os.dup(1)                      # NCCL's version banner goes to stdout: park fd 1 on stderr meanwhile
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    visual = FusedVisual(synthetic.random_vision_state_dict(0, device=dev), device=dev, dtype=torch.float16)
    enc = ZoomEncoder(visual, FusedImageProcessor(min_pixels=3136, max_pixels=16384 * 28 * 28, device=dev))
    g = torch.Generator(device=dev).manual_seed(7)          # same pool on every rank (images are replicated)
    pool = [torch.randint(0, 256, (5000, 5000, 3), generator=g, dtype=torch.uint8, device=dev) for _ in range(args.pool)]

    for cfg_no in args.config:
        if cfg_no == 3:
            nq = 256 if args.full else 64
            tb = synthetic.trajectory_boxes(nq)
            passes = [(f"zoom depth {d + 1}", tb[:, d], np.arange(nq) % args.pool) for d in range(3)]
            what = f"configs[2]: {nq} questions x 3 nested zoom crops (cut_image min 512), one ragged batch per depth"
        elif cfg_no == 4:
            n = 1024 if args.full else 256
            bx, ix = synthetic.mixed_crop_boxes(n, args.pool)
            passes = [("mixed crops", bx, ix)]
            what = f"configs[3]: {n} mixed-size crops (256-2048 px per side) of {args.pool} source images"
        elif cfg_no == 5:
            k = 64 if args.full else 4 * world
            bx = synthetic.maxres_boxes(k, k)
            passes = [("max-res crops", bx, np.arange(2 * k) % args.pool)]
            what = f"configs[4]: {k} crops of 3584x3584 + {k} full 5000x5000 images at max_pixels = 16384*28*28"
        else:
            raise SystemExit(f"unknown config {cfg_no}")

        tot_tokens, tot_ms, fl, k1_total = 0, 0.0, np.zeros(3), 0
        cls_ms = {k: 0.0 for k in KCLASS}
        per_pass = []
        for name, boxes, index in passes:
            # this rank's share: LPT over the FLOP estimate of each crop (identical on every rank)
            from zoomearth_b200 import geometry
            cfg = enc.processor._cfg()
            img_hw = np.array([[5000, 5000]] * len(boxes), np.int32)
            _, _, grid_all = geometry.geometry(cfg, img_hw, np.asarray(boxes, np.float64))
            parts = sharding.partition(sharding.crop_cost(grid_all), world)
            mine = parts[rank]
            my_boxes, my_index = boxes[mine], index[mine]

            k1_bytes = [0]

            def run():
                emb, grid, crop = enc.encode_batched(pool, my_boxes, my_index, args.max_patches)
                gg = grid.numpy()
                k1_bytes[0] = int(((crop[:, 2] - crop[:, 0]).astype(np.int64) * (crop[:, 3] - crop[:, 1]) * 3).sum()
                                  + (gg[:, 1] * gg[:, 2]).sum() * 1176 * 2)
                if world > 1:
                    tokens = enc.tokens_per_crop(grid.numpy())
                    emb, _ = sharding.gather_embeddings(emb, tokens, parts)
                return emb, grid

            import time
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run()                                           # cold pass: builds the plans, sizes the workspaces
            torch.cuda.synchronize()
            cold_ms = (time.perf_counter() - t0) * 1e3
            if world > 1:
                dist.barrier()
            lib.zv_timing_reset()
            lib.zv_timing_enable(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            emb, grid = run()
            e1.record()
            torch.cuda.synchronize()
            lib.zv_timing_enable(0)
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = t.item()
            for k, cid in KCLASS.items():
                t, n = C.c_double(), C.c_int64()
                lib.zv_timing_read(cid, C.byref(t), C.byref(n))
                cls_ms[k] += t.value
            tokens = int((grid_all[:, 1] * grid_all[:, 2]).sum() // 4)
            assert emb.shape[0] == tokens, (emb.shape, tokens)
            assert torch.isfinite(emb[:: max(1, emb.shape[0] // 4096)].float()).all()
            f = np.array(flops(grid_all))
            fl += f
            tot_tokens += tokens
            k1_total += k1_bytes[0]
            tot_ms += ms
            per_pass.append({"pass": name, "crops": len(boxes), "tokens": tokens, "ms": round(ms, 2), "cold_wall_ms": round(cold_ms, 2),
                             "tokens_per_s": round(tokens / ms * 1e3, 1),
                             "tokens_per_crop_min_mean_max": [int((grid_all[:, 1] * grid_all[:, 2]).min() // 4),
                                                              int(tokens / len(boxes)),
                                                              int((grid_all[:, 1] * grid_all[:, 2]).max() // 4)]})
        if rank == 0:
            if stdout_fd is not None:
                sys.stdout.flush()
                C.CDLL(None).fflush(None)       # NCCL's banner may still sit in the C stdio buffer of stdout
                os.dup2(stdout_fd, 1)
                stdout_fd = None
            print(json.dumps({
                "workload": what, "n_gpus": world, "tokens": tot_tokens, "ms": round(tot_ms, 2),
                "tokens_per_s": tot_tokens / tot_ms * 1e3,
                "tower_tflops": float(fl.sum()) / tot_ms / 1e9,
                "flop_split": {"linear": fl[0] / fl.sum(), "window_attn": fl[1] / fl.sum(), "full_attn": fl[2] / fl.sum()},
                "k1_rank0": {"algorithmic_bytes": k1_total, "ms": round(cls_ms["k1_hpass"] + cls_ms["k1_vpass"], 3),
                             "GBps": k1_total / max(1e-9, cls_ms["k1_hpass"] + cls_ms["k1_vpass"]) / 1e6},
                "kernel_ms_rank0": {k: round(v, 2) for k, v in cls_ms.items()},
                "micro_batch_patches": args.max_patches, "passes": per_pass,
                "gather": "nccl ragged all-gather (counts + padded all_gather + permutation)" if world > 1 else "none",
            }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
