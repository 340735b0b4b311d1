"""vLLM plug-in: K1 + the fused tower as vLLM's multimodal encoder for Qwen2.5-VL / ZoomEarth.

What it replaces in the reference: the serving path of ``src/eval/infer_vllm.py`` + ``README.md:98-110``
(``vllm serve PATH_TO_ZOOM_EARTH_MODEL``), whose client round-trips every crop through a base64 **JPEG** data URL
(``infer_vllm.py:126-132``: lossy, so the server never sees the pixels ``infer.py`` feeds the model).

Two pieces, both optional to each other:

* **Server side** - ``register()`` is a ``vllm.general_plugins`` entry point (see INTEGRATION.md for the two lines of
  packaging metadata).  It registers ``ZoomEarthQwen2_5_VLForConditionalGeneration`` for the architecture name
  ``Qwen2_5_VLForConditionalGeneration``: vLLM's own model class with ``self.visual`` swapped for
  ``FusedVllmVisionTransformer`` - same constructor and call surface as vLLM's ``Qwen2_5_VisionTransformer``
  (``forward(x (S, 1176), grid_thw: list[list[int]]) -> (T, out_hidden)``, ``load_weights``, ``dtype`` / ``device`` /
  ``spatial_merge_size`` / ``out_hidden_size``), backed by ``zv_visual_forward``.  The language model, scheduler, KV
  cache and API stay vLLM's.
* **Client side** - ``image_embeds_part(...)`` builds the OpenAI-API content part vLLM accepts under
  ``--enable-mm-embeds`` (``{"type": "image_embeds", "image_embeds": {"image_embeds": b64, "image_grid_thw": b64}}``):
  a client that has the image on a GPU runs crop -> K1 -> tower itself (``ZoomEncoder``) and ships lossless fp16
  embeddings instead of a JPEG; ``png_data_url`` is the lossless drop-in for ``encode_pil_image_to_data_url`` when the
  tower runs in the server.

Nothing here is imported by the rest of the package; vLLM is only imported inside ``register`` / the model factory.
"""
import base64
import io

import numpy as np
import torch
from torch import nn

ARCHITECTURE = "Qwen2_5_VLForConditionalGeneration"


class FusedVllmVisionTransformer(nn.Module):
    """Drop-in for ``vllm.model_executor.models.qwen2_5_vl.Qwen2_5_VisionTransformer`` (same ``__init__`` keywords).

    Holds no ``nn.Parameter``: the checkpoint tensors handed to ``load_weights`` (HF names relative to ``visual.``) are
    packed once into libzoomvit's layout on first use and then dropped."""

    def __init__(self, vision_config, norm_eps=1e-6, quant_config=None, prefix="", dtype=None, device=None):
        super().__init__()
        if quant_config is not None:
            raise NotImplementedError("the fused tower takes 16-bit checkpoints (no quantised vision weights)")
        c = vision_config
        self.vision_config = c
        self.hidden_size, self.num_heads, self.out_hidden_size = c.hidden_size, c.num_heads, c.out_hidden_size
        self.window_size, self.patch_size, self.spatial_merge_size = c.window_size, c.patch_size, c.spatial_merge_size
        self.fullatt_block_indexes = list(c.fullatt_block_indexes)
        self.spatial_merge_unit = self.spatial_merge_size ** 2
        self.norm_eps = norm_eps
        self._dtype = dtype or torch.get_default_dtype()
        if self._dtype not in (torch.float16, torch.bfloat16, torch.float32):
            self._dtype = torch.float16
        self._device = torch.device(device) if device is not None else None
        self._pending = {}
        self._fused = None

    # ---- what vLLM reads
    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        if self._device is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        return self._device

    def load_weights(self, weights):
        """Collects the ``visual.*`` tensors of the checkpoint (HF names, prefix already stripped by vLLM's loader).
        Returns the names it took, as vLLM's ``AutoWeightsLoader`` expects."""
        taken = set()
        for name, w in weights:
            self._pending[name] = w.detach()
            taken.add(name)
        self._fused = None
        return taken

    def _ensure(self):
        if self._fused is None:
            if not self._pending:
                raise RuntimeError("FusedVllmVisionTransformer: no weights loaded")
            from .visual import FusedVisual, _vision_cfg_from_hf
            self._fused = FusedVisual(self._pending, device=self.device, dtype=self._dtype, eps=float(self.norm_eps),
                                      **_vision_cfg_from_hf(self.vision_config))
            self._pending = {}
        return self._fused

    def forward(self, x, grid_thw, *, encoder_metadata=None):
        """x: (S, 1176) patches in HF order (what the HF / fused processor emits); grid_thw: list of [t, h, w]."""
        fv = self._ensure()
        grid = torch.as_tensor(np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3))
        return fv(x.to(self.device), grid)

    # vLLM's encoder CUDA-graph hook asks for per-batch metadata; the fused tower keeps its own plan cache
    def prepare_encoder_metadata(self, grid_thw, *args, **kwargs):
        return {}


def make_model_class():
    """Builds the vLLM model class (imports vLLM)."""
    from vllm.model_executor.models import qwen2_5_vl as q
    from vllm.multimodal import MULTIMODAL_REGISTRY

    @MULTIMODAL_REGISTRY.register_processor(q.Qwen2_5_VLMultiModalProcessor, info=q.Qwen2_5_VLProcessingInfo,
                                            dummy_inputs=q.Qwen2_5_VLDummyInputsBuilder)
    class ZoomEarthQwen2_5_VLForConditionalGeneration(q.Qwen2_5_VLForConditionalGeneration):
        """vLLM's Qwen2.5-VL with the vision tower running on libzoomvit (everything else unchanged)."""

        def __init__(self, *, vllm_config, prefix=""):
            original = q.Qwen2_5_VisionTransformer
            q.Qwen2_5_VisionTransformer = FusedVllmVisionTransformer      # picked up by the parent's constructor
            try:
                super().__init__(vllm_config=vllm_config, prefix=prefix)
            finally:
                q.Qwen2_5_VisionTransformer = original

    return ZoomEarthQwen2_5_VLForConditionalGeneration


def __getattr__(name):                      # lazy: `zoomearth_b200.vllm_plugin:ZoomEarthQwen2_5_VLForConditionalGeneration`
    if name == "ZoomEarthQwen2_5_VLForConditionalGeneration":
        cls = make_model_class()
        globals()[name] = cls
        return cls
    raise AttributeError(name)


def register():
    """``vllm.general_plugins`` entry point: route the Qwen2.5-VL architecture to the fused-tower model class."""
    from vllm import ModelRegistry
    ModelRegistry.register_model(ARCHITECTURE, "zoomearth_b200.vllm_plugin:ZoomEarthQwen2_5_VLForConditionalGeneration")


# ------------------------------------------------------------------------------------------------- client side
def _b64_tensor(t):
    buf = io.BytesIO()
    torch.save(t.detach().cpu().contiguous(), buf)
    return base64.b64encode(buf.getvalue()).decode("utf-8")


def image_embeds_part(embeddings, grid_thw):
    """OpenAI-API content part carrying one image as precomputed tower output (vLLM ``--enable-mm-embeds``): replaces
    ``{"type": "image_url", "image_url": {"url": encode_pil_image_to_data_url(crop)}}`` of infer_vllm.py:126-132,197-199.
    embeddings: (T, out_hidden) from ``ZoomEncoder.encode``; grid_thw: that image's (3,) or (1, 3) grid."""
    g = torch.as_tensor(np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3))
    if g.shape[0] != 1 or int(g[0, 0] * g[0, 1] * g[0, 2]) // 4 != embeddings.shape[0]:
        raise ValueError("one image per part: grid_thw must be a single row matching the number of embeddings")
    return {"type": "image_embeds", "image_embeds": {"image_embeds": _b64_tensor(embeddings), "image_grid_thw": _b64_tensor(g[0])}}


def png_data_url(image_u8):
    """Lossless counterpart of ``encode_pil_image_to_data_url`` (infer_vllm.py:126-132 saves JPEG): (H, W, 3) uint8 ->
    ``data:image/png;base64,...`` so a server-side tower sees the pixels infer.py would have fed it."""
    from PIL import Image
    a = image_u8.cpu().numpy() if isinstance(image_u8, torch.Tensor) else np.asarray(image_u8)
    buf = io.BytesIO()
    Image.fromarray(np.ascontiguousarray(a)).save(buf, format="PNG")
    return "data:image/png;base64," + base64.b64encode(buf.getvalue()).decode("utf-8")
