"""On-wire outputs of the sampling script, fed from the decoder's uint8 epilogue.

scripts/sample_diffusion.py writes two things per run:
  * `<nplog>/<N>x<H>x<W>x3-samples.npz` — `np.savez(path, all_img)` of every sample as uint8 NHWC, made by `custom_to_np`
    (:115-121, :293-301; array name `arr_0`, truncated to n_samples);
  * `<logdir>/<key>/<file_name or key_%06d>.png` — one PNG per image through `custom_to_pil` (:103-108, :306-334).
Here both start from uint8 NHWC device tensors (`FridoDiffusion.decode_first_stage_uint8`, modes "np" / "pil"): the bytes
are copied to PINNED host buffers on a side stream (the sampler keeps the main stream), PNG encoding runs on worker
threads straight from the pinned memory, and the .npz is assembled from the same buffers.  No fp32 image crosses PCIe.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch


class SampleWriter:
    def __init__(self, logdir=None, nplog=None, n_samples=50000, key="sample", workers=4, pinned_slots=2):
        self.logdir, self.nplog, self.n_samples, self.key = logdir, nplog, n_samples, key
        self.n_saved = 0
        self._np_chunks = []
        self._slots = [None] * max(1, pinned_slots)  # (np pinned, pil pinned, event, pending futures)
        self._k = 0
        self._pool = ThreadPoolExecutor(max_workers=workers) if logdir is not None else None
        self._stream = None
        if logdir is not None:
            os.makedirs(os.path.join(logdir, key), exist_ok=True)
        if nplog is not None:
            os.makedirs(nplog, exist_ok=True)

    # ------------------------------------------------------------------
    def _pinned_like(self, t, old):
        if old is not None and old.shape == t.shape:
            return old
        return torch.empty(tuple(t.shape), dtype=torch.uint8, pin_memory=t.is_cuda)

    def add_batch(self, u8_np, u8_pil=None, file_names=None):
        """u8_np: uint8 [B,H,W,3] in `custom_to_np` format (device or host tensor); u8_pil: the same images in `custom_to_pil`
        format for the PNGs (defaults to u8_np when the two roundings are not needed separately).  Returns immediately
        after enqueueing the device->host copies; `finish()` (or the next reuse of the pinned slot) waits for them."""
        slot = self._k % len(self._slots)
        self._k += 1
        old = self._slots[slot]
        if old is not None:
            self._retire(old)
        want_png = self.logdir is not None
        src_pil = u8_pil if u8_pil is not None else u8_np
        h_np = self._pinned_like(u8_np, old[0] if old else None)
        h_pil = self._pinned_like(src_pil, old[1] if old else None) if (want_png and u8_pil is not None) else None
        ev = None
        if u8_np.is_cuda:
            dev = u8_np.device
            if self._stream is None:
                self._stream = torch.cuda.Stream(dev)
            produced = torch.cuda.current_stream(dev).record_event()
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(produced)
                h_np.copy_(u8_np, non_blocking=True)
                if h_pil is not None:
                    h_pil.copy_(u8_pil, non_blocking=True)
                ev = self._stream.record_event()
            u8_np.record_stream(self._stream)
            if u8_pil is not None:
                u8_pil.record_stream(self._stream)
        else:
            h_np.copy_(u8_np)
            if h_pil is not None:
                h_pil.copy_(u8_pil)
        names = list(file_names) if file_names is not None else None
        first = self.n_saved
        self.n_saved += int(u8_np.shape[0])
        self._slots[slot] = [h_np, h_pil, ev, names, first, False]

    def _retire(self, s):
        """Wait for a slot's copies, take its samples into the .npz list, encode its PNGs."""
        h_np, h_pil, ev, names, first, done = s
        if done:
            return
        if ev is not None:
            ev.synchronize()
        arr = h_np.numpy()
        if self.nplog is not None:
            self._np_chunks.append(arr.copy())
        if self._pool is not None:
            img = (h_pil if h_pil is not None else h_np).numpy()
            futs = []
            for i in range(arr.shape[0]):
                if names is not None:
                    fn = "{}.png".format(str(names[i]).split(".")[0])  # sample_diffusion.py:320-321
                else:
                    fn = f"{self.key}_{first + i:06}.png"              # :322-323
                futs.append(self._pool.submit(_save_png, img[i].copy(), os.path.join(self.logdir, self.key, fn)))
            for f in futs:
                f.result()
        s[5] = True

    def finish(self):
        """Flush; returns the path of the .npz (None without `nplog`)."""
        order = sorted((s for s in self._slots if s is not None), key=lambda s: s[4])
        for s in order:
            self._retire(s)
        if self._pool is not None:
            self._pool.shutdown(wait=True)
        if self.nplog is None:
            return None
        all_img = np.concatenate(self._np_chunks, axis=0)[: self.n_samples] if self._np_chunks else np.zeros((0,), np.uint8)
        shape_str = "x".join(str(x) for x in all_img.shape)
        path = os.path.join(self.nplog, f"{shape_str}-samples.npz")  # sample_diffusion.py:297-301
        np.savez(path, all_img)
        return path


def _save_png(hwc_u8, path):
    from PIL import Image

    im = Image.fromarray(hwc_u8 if hwc_u8.shape[-1] != 1 else hwc_u8[..., 0])
    if im.mode != "RGB":
        im = im.convert("RGB")  # sample_diffusion.py:110-111
    im.save(path)
