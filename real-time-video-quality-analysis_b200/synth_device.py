"""Per-frame deterministic synthetic workloads on a torch device (bench.py; a WORKLOAD definition, not
part of the measured path).

Same ingredients as the host recipe of ``synth.py`` / SURVEY.md 8(d) -- a blurred-noise background panned
3 px/frame in x and 2 px/frame in y, a moving white disc, a moving red rectangle, 0..7 of additive
per-pixel noise -- but every frame is a pure function of ``(seed, frame index)``: a rank can generate
any frame range of a long clip (BASELINE.json config 4: one 4K clip of 3600 frames cut into frame ranges;
config 5: 64 clips) without generating what comes before it, and every rank count sees the same frames.
The background is periodic (circular blur, ``torch.roll``), so a clip of any length needs one texture.

``pairs(first, n)`` returns what the reference's flow holds for a clip: the SOURCE as yuv420p planes (BT.601
limited range of the synthetic BGR frame) and its ENCODE as yuv420p planes (3x3 blur + {-2..2} noise per
plane: the stand-in for the CRF-23 re-encode of config 3; there is no x264 in the image).  The BGR frames
the complexity metrics see are the decode of the encode (yuv420p -> BGR, done inside the library).
"""
from __future__ import annotations

import numpy as np


class DeviceClipSynth:
    def __init__(self, h: int, w: int, seed: int, device):
        import torch
        if (h | w) & 1:
            raise ValueError("yuv420p clips need even frame sizes")
        self.h, self.w, self.seed, self.dev = int(h), int(w), int(seed), torch.device(device)
        self.torch = torch
        rng = np.random.default_rng(1_000_000 + seed)
        raw = torch.from_numpy(rng.integers(0, 256, (3, h, w), dtype=np.uint8)).to(self.dev).float()
        # circular Gaussian blur (sigma 2) -> periodic texture, stretched to 0..255
        r = 6
        x = torch.arange(-r, r + 1, device=self.dev, dtype=torch.float32)
        k = torch.exp(-(x * x) / 8.0)
        k = k / k.sum()
        t = torch.nn.functional.pad(raw[None], (r, r, r, r), mode="circular")
        t = torch.nn.functional.conv2d(t, k.view(1, 1, 1, -1).repeat(3, 1, 1, 1), groups=3)
        t = torch.nn.functional.conv2d(t, k.view(1, 1, -1, 1).repeat(3, 1, 1, 1), groups=3)[0]
        lo, hi = t.min(), t.max()
        t = ((t - lo) * (255.0 / torch.clamp(hi - lo, min=1e-6))).round().clamp(0, 255).to(torch.uint8)
        self.tex = t.permute(1, 2, 0).contiguous()                     # [h, w, 3] B,G,R
        self.yy = torch.arange(h, device=self.dev, dtype=torch.float32).view(h, 1)
        self.xx = torch.arange(w, device=self.dev, dtype=torch.float32).view(1, w)
        self.gen = torch.Generator(device=self.dev)

    # ---------------------------------------------------------------- one frame
    def bgr(self, i: int):
        """Source frame i, [h, w, 3] uint8 on the device."""
        torch, h, w = self.torch, self.h, self.w
        f = torch.roll(self.tex, shifts=(-((2 * i) % h), -((3 * i) % w)), dims=(0, 1)).clone()
        cx, cy = (w / 4.0 + 5 * i) % w, (h / 3.0 + 2 * i) % h
        disc = (self.xx - cx) ** 2 + (self.yy - cy) ** 2 <= (h / 12.0) ** 2
        f[disc] = 255
        x0, x1 = int(w / 2), min(int(w / 2 + w / 6), w)
        y0 = int(h / 5 + 3 * i) % h
        y1 = min(y0 + int(h / 4), h)
        f[y0:y1, x0:x1, 0] = 0
        f[y0:y1, x0:x1, 1] = 0
        f[y0:y1, x0:x1, 2] = 255
        self.gen.manual_seed(self.seed * 1_000_003 + 2 * i)
        noise = torch.randint(0, 8, (h, w, 3), generator=self.gen, device=self.dev, dtype=torch.int16)
        return (f.to(torch.int16) + noise).clamp_(max=255).to(torch.uint8)

    def _blur3(self, p):
        torch = self.torch
        q = torch.nn.functional.pad(p.to(torch.int16)[None, None], (1, 1, 1, 1), mode="replicate")[0, 0]
        hs = q[:, :-2] + 2 * q[:, 1:-1] + q[:, 2:]
        return (hs[:-2] + 2 * hs[1:-1] + hs[2:] + 8) >> 4

    def pair(self, i: int):
        """(source planes, encode planes) of frame i: 2 x (Y [h,w], U [h/2,w/2], V) uint8."""
        torch = self.torch
        x = self.bgr(i).to(torch.int32)
        b, g, r = x[..., 0], x[..., 1], x[..., 2]
        y = ((66 * r + 129 * g + 25 * b + 128) >> 8) + 16
        u = ((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128
        v = ((112 * r - 94 * g - 18 * b + 128) >> 8) + 128

        def sub(p):
            return (p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + 2) >> 2

        ref = [y.clamp(0, 255).to(torch.uint8), sub(u).clamp(0, 255).to(torch.uint8), sub(v).clamp(0, 255).to(torch.uint8)]
        self.gen.manual_seed(self.seed * 1_000_003 + 2 * i + 1)
        enc = []
        for p in ref:
            n = torch.randint(-2, 3, p.shape, generator=self.gen, device=self.dev, dtype=torch.int16)
            enc.append((self._blur3(p) + n).clamp_(0, 255).to(torch.uint8))
        return ref, enc

    # ---------------------------------------------------------------- ranges
    def pairs(self, first: int, n: int, out=None):
        """Frames first .. first+n-1 as plane stacks: (source [Y,U,V], encode [Y,U,V]), Y [n,h,w], U/V [n,h/2,w/2]."""
        torch, h, w = self.torch, self.h, self.w
        if out is None:
            mk = lambda hh, ww: torch.empty((n, hh, ww), dtype=torch.uint8, device=self.dev)
            out = ([mk(h, w), mk(h // 2, w // 2), mk(h // 2, w // 2)], [mk(h, w), mk(h // 2, w // 2), mk(h // 2, w // 2)])
        for j in range(n):
            ref, enc = self.pair(first + j)
            for c in range(3):
                out[0][c][j] = ref[c]
                out[1][c][j] = enc[c]
        return out
