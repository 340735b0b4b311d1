"""Tower parity (GPU, through the C ABI) vs the fp32 oracle: cosine >= 0.999 and max|a-b|/max|b| <= 1e-2
(the tolerances BASELINE.json's north_star states) on the shipped operand dtype (fp16 - the one bench.py times), at
full depth on the bench shape, a ragged batch and zoom crops; integer artefacts bit-exact."""
import numpy as np
import pytest
import torch

from oracle import processor as OP, resample as OR, tower as OT

pytestmark = pytest.mark.gpu


def _metrics(got, ref):
    """(cosine of the flattened tensors, max|a-b| / max|b|), accumulated in float64 (fp32 sums over 1e8 elements drift)."""
    got, ref = got.detach().double().cpu().flatten(), ref.detach().double().cpu().flatten()
    cos = (torch.dot(got, ref) / (got.norm() * ref.norm())).item()
    return cos, ((got - ref).abs().max() / ref.abs().max()).item()


def _fused(sd, cfg, cuda, dtype=torch.float32, operand_dtype=None):
    from zoomearth_b200 import FusedVisual
    return FusedVisual(sd, device=cuda, dtype=dtype, operand_dtype=operand_dtype, depth=cfg["depth"],
                       fullatt=list(cfg["fullatt"]))


@pytest.mark.parametrize("operand", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("grid", [[[1, 2, 2]], [[1, 8, 8]], [[1, 26, 36], [1, 8, 8], [1, 18, 34], [1, 2, 2]], [[1, 36, 36]],
                                  [[1, 2, 200]]])
def test_tower_small_depth_vs_oracle(cuda, grid, operand):
    cfg = OT.small_cfg(depth=3, fullatt=(1,))
    sd = OT.make_weights(1, cfg)
    grid = np.array(grid)
    S = int((grid[:, 1] * grid[:, 2]).sum())
    pv = torch.randn(S, 1176, generator=torch.Generator().manual_seed(S))
    fv = _fused(sd, cfg, cuda, operand_dtype=operand)
    out, hidden = fv(pv.to(cuda), torch.from_numpy(grid), return_hidden=True)
    ref, ref_hidden = OT.forward(sd, pv, grid, cfg, return_hidden=True)
    ch, mh = _metrics(hidden, ref_hidden)
    co, mo = _metrics(out, ref)
    assert ch >= 0.999 and mh <= 1e-2, f"hidden: cos {ch} maxrel {mh}"
    assert co >= 0.999 and mo <= 1e-2, f"merged: cos {co} maxrel {mo}"
    if operand == torch.bfloat16:
        # same precision policy on the CPU (bf16 operands, fp32 accumulate / residual) must agree much tighter
        emu = OT.forward(sd, pv, grid, cfg, emulate_bf16=True)
        ce, me = _metrics(out, emu)
        assert me <= 8e-3, f"vs bf16-operand emulation: cos {ce} maxrel {me}"
    else:
        assert mo <= 2e-3 and mh <= 2e-3, f"fp16 operands at depth 3: merged {mo}, hidden {mh}"


def _pixels(S, seed):
    """Patch rows with the statistics of real processor output: uniform uint8 noise through rescale + normalize."""
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (S, 1176), generator=g).float()
    return (u8 / 255.0 - 0.45) / 0.27


# Full depth (32 blocks, full attention at 7/15/23/31), the shipped operand dtype (fp16 - what bench.py times).  The
# checker is the fp32 oracle run on the GPU with TF32 off (oracle.tower.forward(device=...)); the first case also
# pins that GPU run to the CPU run of the same oracle.
FULL_DEPTH_CASES = [
    ("zoom crop 504x504 + small image, randn patches", [[1, 36, 36], [1, 10, 14]], "randn", 7),
    ("global view 980x980: the bench shape, one image", [[1, 70, 70]], "pixels", 70),
    ("ragged batch of six crops incl. a 64x92 zoom crop and a 4x200 strip",
     [[1, 26, 36], [1, 8, 8], [1, 18, 34], [1, 2, 2], [1, 64, 92], [1, 4, 200]], "pixels", 11),
    ("two global views + two zoom crops", [[1, 70, 70], [1, 36, 36], [1, 70, 70], [1, 44, 58]], "randn", 5),
]


@pytest.mark.parametrize("name,grid,kind,seed", FULL_DEPTH_CASES, ids=[c[0].split(":")[0].split(",")[0] for c in FULL_DEPTH_CASES])
def test_tower_full_depth_timed_dtype_vs_oracle(cuda, full_sd, full_sd_cuda, full_visual, name, grid, kind, seed):
    grid = np.array(grid)
    S = int((grid[:, 1] * grid[:, 2]).sum())
    pv = torch.randn(S, 1176, generator=torch.Generator().manual_seed(seed)) if kind == "randn" else _pixels(S, seed)
    ref = OT.forward(full_sd_cuda, pv, grid, device=cuda)
    if S < 2000:                                    # pin the GPU run of the oracle to its CPU run
        ref_cpu = OT.forward(full_sd, pv, grid)
        c, m = _metrics(ref, ref_cpu)
        assert m <= 2e-5, f"fp32 oracle on the GPU differs from the CPU run: maxrel {m}"
    out = full_visual(pv.to(cuda), torch.from_numpy(grid))
    assert out.shape == (S // 4, 2048)
    cos, maxrel = _metrics(out, ref)
    print(f"PARITY full tower fp16 operands vs fp32 oracle [{name}]: S={S} cosine {cos:.6f}, max rel err {maxrel:.3e}")
    assert cos >= 0.999 and maxrel <= 1e-2, f"{name}: cos {cos} maxrel {maxrel}"
    assert maxrel <= 5e-3, f"fp16 operands should sit near 1.5e-3, got {maxrel}"


def test_tower_full_depth_fp16_output_dtype(cuda, full_sd, full_sd_cuda):
    """The dtype the reference's eval loop runs in (infer.py:149) end to end: fp16 operands AND fp16 embeddings."""
    grid = np.array([[1, 36, 36], [1, 10, 14]])
    pv = torch.randn(int((grid[:, 1] * grid[:, 2]).sum()), 1176, generator=torch.Generator().manual_seed(7))
    fv = _fused(full_sd, OT.CFG, cuda, dtype=torch.float16)
    out = fv(pv.to(cuda), torch.from_numpy(grid))
    assert out.dtype == torch.float16 and fv.operand_dtype == torch.float16
    cos, maxrel = _metrics(out, OT.forward(full_sd_cuda, pv, grid, device=cuda))
    assert cos >= 0.999 and maxrel <= 1e-2, f"cos {cos} maxrel {maxrel}"


def test_tower_full_depth_bf16_operands_opt_in(cuda, full_sd, full_sd_cuda):
    """bf16 operands are an OPT-IN (operand_dtype=torch.bfloat16), not the shipped or benchmarked configuration: with
    8 mantissa bits the full-depth max-rel error sits above the 1e-2 north_star states (measured 1.15e-2 here; the
    CPU emulation of the identical policy 1.26e-2; HF's own bf16 tower 2.2e-2).  What is asserted for this
    mode is cosine >= 0.999 and 'no worse than 1.25x the emulated policy'; the stated tolerance is asserted on the
    shipped dtype in the tests above."""
    grid = np.array([[1, 36, 36], [1, 10, 14]])
    pv = torch.randn(int((grid[:, 1] * grid[:, 2]).sum()), 1176, generator=torch.Generator().manual_seed(7))
    ref = OT.forward(full_sd_cuda, pv, grid, device=cuda)
    emu = OT.forward(full_sd_cuda, pv, grid, device=cuda, emulate_bf16=True)
    fv = _fused(full_sd, OT.CFG, cuda, dtype=torch.bfloat16, operand_dtype=torch.bfloat16)
    out = fv(pv.to(cuda), torch.from_numpy(grid))
    assert out.dtype == torch.bfloat16
    cos, maxrel = _metrics(out, ref)
    _, emurel = _metrics(emu, ref)
    print(f"PARITY full tower, bf16 operands (opt-in) vs fp32 oracle: cosine {cos:.6f}, max rel err {maxrel:.3e}; "
          f"emulated bf16-operand policy {emurel:.3e}")
    assert cos >= 0.999 and maxrel <= 1.25 * emurel, f"cos {cos} maxrel {maxrel} (emulated policy {emurel})"


def test_plan_artefacts_bitexact_vs_hf(cuda):
    """window_index, cu_window_seqlens, cu_seqlens and rotary ids equal the live HF tower's own functions."""
    from oracle import hf_live
    from zoomearth_b200 import Plan, _lib
    cfg = OT.small_cfg(depth=1, fullatt=())
    m = hf_live.hf_tower(None, cfg)
    grid = torch.tensor([[1, 70, 70], [1, 26, 36], [1, 64, 92], [1, 4, 200]])
    p = Plan(_lib.default_cfg(), grid.numpy())
    w, cu = m.get_window_index(grid)
    assert np.array_equal(p.window_index, w.numpy())
    assert p.cu_window_seqlens_raw.tolist() == cu
    assert p.cu_window_seqlens.tolist() == torch.unique_consecutive(torch.tensor(cu)).tolist()
    assert np.array_equal(p.pos_ids, OT.rot_pos_ids(grid.numpy()))


def test_zoom_step_end_to_end(cuda):
    """crop -> K1 (bf16, window order) -> tower == oracle crop/resize/patchify -> fp32 oracle tower."""
    from zoomearth_b200 import FusedImageProcessor, ZoomEncoder
    cfg = OT.small_cfg(depth=4, fullatt=(3,))
    sd = OT.make_weights(2, cfg)
    img = np.random.default_rng(3).integers(0, 256, (1500, 2000, 3), dtype=np.uint8)
    boxes = [(100, 200, 1300, 1100), (900.5, 700.2, 1100.9, 800.0)]
    fv = _fused(sd, cfg, cuda, operand_dtype=torch.float16)
    enc = ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=401408, device=cuda))
    dev = enc.upload(img)
    emb, grid, crop = enc.encode([dev], boxes, image_index=[0, 0])
    rows, grids = [], []
    for b in boxes:
        box, pv, g = OP.zoom_step_u8(img, b, 512, 3136, 401408)
        rows.append(pv)
        grids.append(g)
    rgrid = np.concatenate(grids, 0)
    assert grid.tolist() == rgrid.tolist()
    ref = OT.forward(sd, torch.from_numpy(np.concatenate(rows, 0)), rgrid, cfg)
    cos, maxrel = _metrics(emb, ref)
    assert cos >= 0.999 and maxrel <= 1e-2, f"cos {cos} maxrel {maxrel}"
    assert enc.last_launches > 0


def test_encode_host_pipeline_matches_resident_path(cuda):
    """The chunked host pipeline (H2D / compute / D2H on three streams) returns what the resident path does.  Not
    bitwise: the full-attention kernel aligns its K/V tiles to 8 rows of the batch, so a segment's tile split (and
    with it the rounding of the online softmax) depends on where the segment starts in the batch."""
    from zoomearth_b200 import FusedImageProcessor, ZoomEncoder
    cfg = OT.small_cfg(depth=2, fullatt=(1,))
    sd = OT.make_weights(4, cfg)
    fv = _fused(sd, cfg, cuda, dtype=torch.float16)
    enc = ZoomEncoder(fv, FusedImageProcessor(min_pixels=3136, max_pixels=100352, device=cuda))
    rng = np.random.default_rng(9)
    host = [torch.from_numpy(rng.integers(0, 256, (400 + 40 * i, 500, 3), dtype=np.uint8)).pin_memory() for i in range(5)]
    out, grid = enc.encode_host(host, chunk=2)
    ref, rgrid, _ = enc.encode([h.to(cuda) for h in host], None)
    assert grid.tolist() == rgrid.tolist()
    assert out.shape == ref.shape
    cos, maxrel = _metrics(out, ref)
    assert cos >= 0.9995 and maxrel <= 1e-2, f"cos {cos} maxrel {maxrel}"
