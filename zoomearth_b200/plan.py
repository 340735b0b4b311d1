"""Per-batch plan: the integer bookkeeping of the tower, computed by libzoomvit (zv_plan_*).

Mirrors HF modeling_qwen2_5_vl.py: rot_pos_emb ids (:382-401), get_window_index (:411-451), cu_window_seqlens
after unique_consecutive (:476) and cu_seqlens (:488-496).  The NumPy views are copies, safe to keep.
"""
import ctypes as C

import numpy as np

from . import _lib


class Plan:
    def __init__(self, cfg, grid_thw):
        self.cfg = cfg
        g = np.ascontiguousarray(np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3))
        self.grid_thw = g
        h = C.c_void_p()
        _lib.check(_lib.lib().zv_plan_create(C.byref(cfg), g.shape[0], g.ctypes.data, C.byref(h)))
        self._h = h
        self.num_patches = _lib.lib().zv_plan_num_patches(h)
        self.num_tokens = _lib.lib().zv_plan_num_tokens(h)
        self.device_bytes = _lib.lib().zv_plan_device_bytes(h)
        self._dev = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().zv_plan_free(h)
            except Exception:           # interpreter shutdown: the module globals are already gone
                pass

    @property
    def handle(self):
        return self._h

    def _arr(self, ptr, n, dtype):
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

    @property
    def window_index(self):
        return self._arr(_lib.lib().zv_plan_window_index(self._h), self.num_tokens, np.int64)

    @property
    def reverse_index(self):
        return self._arr(_lib.lib().zv_plan_reverse_index(self._h), self.num_tokens, np.int64)

    def _cu(self, fn):
        n = C.c_int32()
        p = fn(self._h, C.byref(n))
        return self._arr(p, n.value, np.int32)

    @property
    def cu_window_seqlens(self):
        return self._cu(_lib.lib().zv_plan_cu_window)

    @property
    def cu_window_seqlens_raw(self):
        return self._cu(_lib.lib().zv_plan_cu_window_raw)

    @property
    def cu_seqlens(self):
        return self._cu(_lib.lib().zv_plan_cu_full)

    @property
    def pos_ids(self):
        return self._arr(_lib.lib().zv_plan_pos_ids(self._h), self.num_patches * 2, np.int32).reshape(-1, 2)

    def device_tables(self, device, stream):
        """Uploads (once) the device-side tables into a torch buffer and returns it."""
        import torch
        if self._dev is None or self._dev.device != device:
            buf = torch.empty(self.device_bytes, dtype=torch.uint8, device=device)
            _lib.check(_lib.lib().zv_plan_upload(self._h, buf.data_ptr(), buf.numel(), stream))
            self._dev = buf
        return self._dev
