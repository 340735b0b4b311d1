"""Decode-once ingest: image files -> resident uint8 device tensors, keyed by path.

The reference opens and decodes the same TIFF/PNG/JPEG twice per question (``src/eval/infer.py:215,237``,
``src/demo.py:131,139``; the GRPO rollout again, ``grpo_trainer.py:530,602``) and ships float pixel_values over PCIe.
``ImageStore`` decodes every file once - with Pillow, the reference's own decoder, so the pixels are bit-identical
(``Image.open(path).convert("RGB")``) - on a small thread pool (Pillow releases the GIL while decoding), stages the
uint8 pixels in pinned memory and uploads them on a side stream, so the decode and the H2D copy of question i+1 run
under the GPU work of question i.  Entries are kept under a byte budget (LRU) and dropped explicitly at the end of a
question; a changed file (mtime / size) is decoded again.

``device="cpu"`` keeps the tensors on the host (no pinning, no streams): the cache / prefetch logic is what the CPU
tests exercise; the product path uses a CUDA device.
"""
import os
import threading
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch


def decode_rgb_u8(path):
    """``np.asarray(Image.open(path).convert("RGB"))`` - the decode of infer.py:215 - as a contiguous (H, W, 3) uint8 array."""
    from PIL import Image
    with Image.open(path) as im:
        return np.ascontiguousarray(np.asarray(im.convert("RGB")))


class ImageStore:
    def __init__(self, device=None, budget_bytes=32 << 30, workers=4, decoder=decode_rgb_u8):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("ImageStore needs a CUDA device (pass device='cpu' only for host-side tests)")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.budget_bytes = int(budget_bytes)
        self.decoder = decoder
        self._pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="zv-decode")
        self._lock = threading.Lock()
        self._entries = OrderedDict()            # key -> dict(sig, future | tensor, event, bytes)
        self._bytes = 0
        self._copy_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.decodes = 0                          # files actually decoded (the tests count these)
        self.hits = 0

    @staticmethod
    def _sig(path):
        st = os.stat(path)
        return (st.st_mtime_ns, st.st_size)

    def _load(self, path):
        """Worker thread: decode, pin, enqueue the upload on the copy stream.  Returns (tensor, ready event or None)."""
        arr = self.decoder(path)
        with self._lock:
            self.decodes += 1
        t = torch.from_numpy(arr)
        if self.device.type != "cuda":
            return t, None
        from .processor import upload_u8
        d = upload_u8(t, self.device, stream=self._copy_stream)       # pinned, async, rows padded when 3 W is not a multiple of 4
        ev = torch.cuda.Event()
        ev.record(self._copy_stream)
        return d, ev

    def prefetch(self, paths):
        """Schedule decode + upload of these files (no-op for resident ones).  Returns immediately."""
        for p in paths:
            self._entry(p)

    def _entry(self, path):
        key = os.path.abspath(path)
        sig = self._sig(key)
        with self._lock:
            e = self._entries.get(key)
            if e is not None and e["sig"] == sig:
                self._entries.move_to_end(key)
                return e
            if e is not None:
                self._bytes -= e.get("bytes", 0)
            e = {"sig": sig, "future": self._pool.submit(self._load, key), "tensor": None, "event": None, "bytes": 0}
            self._entries[key] = e
            return e

    def get(self, path):
        """The (H, W, 3) uint8 tensor of this file on the store's device; decodes at most once per file version.  The
        caller's current CUDA stream is made to wait for the upload (no host synchronisation)."""
        e = self._entry(path)
        fresh = e["tensor"] is None
        if fresh:
            t, ev = e["future"].result()
            with self._lock:
                if e["tensor"] is None:
                    e["tensor"], e["event"], e["bytes"] = t, ev, t.numel()
                    e["future"] = None
                    self._bytes += e["bytes"]
                    self._evict(keep=e)
        else:
            with self._lock:
                self.hits += 1
        if e["event"] is not None:
            torch.cuda.current_stream(self.device).wait_event(e["event"])
        return e["tensor"]

    def _evict(self, keep):
        while self._bytes > self.budget_bytes and len(self._entries) > 1:
            k, e = next(iter(self._entries.items()))
            if e is keep or e["tensor"] is None:
                self._entries.move_to_end(k)
                if all(v is keep or v["tensor"] is None for v in self._entries.values()):
                    break
                continue
            self._entries.pop(k)
            self._bytes -= e["bytes"]

    def drop(self, path):
        key = os.path.abspath(path)
        with self._lock:
            e = self._entries.pop(key, None)
            if e is not None:
                self._bytes -= e.get("bytes", 0)

    def resident_bytes(self):
        return self._bytes

    def __contains__(self, path):
        return os.path.abspath(path) in self._entries

    def close(self):
        self._pool.shutdown(wait=True)
