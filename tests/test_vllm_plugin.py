"""CPU: the vLLM plug-in registers, mirrors the surface of vLLM's own vision transformer, and its client
helpers produce what vLLM's multimodal input parser accepts.  (The tower itself is checked on the GPU in
tests/test_gpu_ingest_plugin.py.)"""
import inspect

import numpy as np
import pytest
import torch

vllm = pytest.importorskip("vllm")


def test_register_routes_the_architecture_to_the_fused_model_class():
    from vllm import ModelRegistry
    from zoomearth_b200 import vllm_plugin as P
    P.register()
    assert P.ARCHITECTURE in ModelRegistry.get_supported_archs()
    cls = P.ZoomEarthQwen2_5_VLForConditionalGeneration
    from vllm.model_executor.models import qwen2_5_vl as q
    assert issubclass(cls, q.Qwen2_5_VLForConditionalGeneration)
    assert q.Qwen2_5_VisionTransformer is not P.FusedVllmVisionTransformer      # the swap is scoped to the constructor


def test_adapter_mirrors_vllm_vision_transformer_surface():
    from vllm.model_executor.models import qwen2_5_vl as q
    from zoomearth_b200.vllm_plugin import FusedVllmVisionTransformer as F
    theirs = inspect.signature(q.Qwen2_5_VisionTransformer.__init__).parameters
    ours = inspect.signature(F.__init__).parameters
    assert all(k in ours for k in theirs), (list(theirs), list(ours))
    tf = inspect.signature(q.Qwen2_5_VisionTransformer.forward).parameters
    of = inspect.signature(F.forward).parameters
    assert list(tf)[:3] == list(of)[:3] == ["self", "x", "grid_thw"] and "encoder_metadata" in of
    for attr in ("dtype", "device", "load_weights", "prepare_encoder_metadata"):
        assert hasattr(F, attr)
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLVisionConfig
    m = F(Qwen2_5_VLVisionConfig(), dtype=torch.float16)
    assert (m.spatial_merge_size, m.out_hidden_size, m.hidden_size, m.patch_size, m.window_size) == (2, 3584, 1280, 14, 112) \
        or m.spatial_merge_size == 2
    taken = m.load_weights([("blocks.0.attn.qkv.weight", torch.zeros(3, 3)), ("merger.ln_q.weight", torch.zeros(3))])
    assert taken == {"blocks.0.attn.qkv.weight", "merger.ln_q.weight"} and len(list(m.parameters())) == 0


def test_client_parts_are_what_vllm_parses():
    from vllm.multimodal.media.image import ImageEmbeddingMediaIO
    from zoomearth_b200.vllm_plugin import image_embeds_part, png_data_url
    emb = torch.randn(6, 2048).half()
    part = image_embeds_part(emb, [[1, 4, 6]])
    assert part["type"] == "image_embeds" and set(part["image_embeds"]) == {"image_embeds", "image_grid_thw"}
    io_ = ImageEmbeddingMediaIO()
    assert torch.equal(io_.load_base64("", part["image_embeds"]["image_embeds"]), emb)
    assert io_.load_base64("", part["image_embeds"]["image_grid_thw"]).tolist() == [1, 4, 6]
    with pytest.raises(ValueError):
        image_embeds_part(emb, [[1, 4, 4]])
    # lossless pixels (the reference's client saves JPEG, infer_vllm.py:126-132)
    import base64, io
    from PIL import Image
    img = np.random.default_rng(1).integers(0, 256, (40, 50, 3), dtype=np.uint8)
    url = png_data_url(img)
    assert url.startswith("data:image/png;base64,")
    back = np.asarray(Image.open(io.BytesIO(base64.b64decode(url.split(",", 1)[1]))).convert("RGB"))
    assert np.array_equal(back, img)
