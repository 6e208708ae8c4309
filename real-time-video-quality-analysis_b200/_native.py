"""ctypes binding of libvqa_b200.so (include/vqa_b200.h).

There is NO CPU fallback: if the shared object is missing or no sm_100 GPU is present the
binding raises.  PyTorch is used only for device memory / streams (CUDA tensors are accepted
as inputs and the context runs on torch's current stream).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libvqa_b200.so")

M_HIST, M_COLOR, M_EDGE, M_DCT, M_ORB, M_MOTION, M_TDCT = (1 << i for i in range(7))
M_ALL = 0x7F

EXPORTS = [
    "vqa_abi_version", "vqa_init", "vqa_destroy", "vqa_last_error", "vqa_set_stream", "vqa_sync",
    "vqa_kernel_launches", "vqa_stage_ms", "vqa_reset_timers", "vqa_complexity_frames",
    "vqa_psnr_ssim_planar", "vqa_analyze_clip", "vqa_framerate_series", "vqa_ewm_partial", "vqa_debug_gray",
    "vqa_debug_resize", "vqa_debug_hist", "vqa_debug_orb", "vqa_kernel_profile", "vqa_kernel_report", "vqa_debug_canny", "vqa_debug_flow", "vqa_debug_dct",
    "vqa_orb_default_cfg", "vqa_orb_describe", "vqa_orb_detect", "vqa_debug_orb_pyramid", "vqa_debug_exact_taps",
    "vqa_analyze_clip_yuv420", "vqa_debug_yuv2bgr",
    "vqa_comm_unique_id", "vqa_comm_init", "vqa_comm_destroy", "vqa_clip_reduce", "vqa_comm_halo_exchange",
]
COMM_ID_BYTES = 128
ABI_VERSION = 3


class VqaError(RuntimeError):
    pass


class Cfg(C.Structure):
    _fields_ = [("resize_width", C.c_int32), ("resize_height", C.c_int32),
                ("metrics_mask", C.c_uint32), ("dct_impl", C.c_int32),
                ("orb_width", C.c_int32), ("orb_height", C.c_int32)]


class OrbCfg(C.Structure):
    """cv2.ORB_create(nfeatures, scaleFactor, nlevels, edgeThreshold, fastThreshold=...) (vqa_orb_cfg)."""
    _fields_ = [("nfeatures", C.c_int32), ("nlevels", C.c_int32), ("edge_threshold", C.c_int32),
                ("fast_threshold", C.c_int32), ("scale_factor", C.c_float)]


def orb_cfg(nfeatures=500, scale_factor=1.2, nlevels=8, edge_threshold=31, fast_threshold=20) -> OrbCfg:
    return OrbCfg(int(nfeatures), int(nlevels), int(edge_threshold), int(fast_threshold), float(scale_factor))


KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("lx", "<i4"),
                           ("ly", "<i4"), ("fast_score", "<i4")], align=True)
ORB_MAX_LEVELS = 16


FRAME_DTYPE = np.dtype([("hist_entropy", "<f4"), ("color_entropy", "<f4"), ("dct_energy", "<f4"),
                        ("motion", "<f4"), ("temporal_dct", "<f4"), ("orb_count", "<i4"),
                        ("edge_count", "<i8"), ("gray_sq_sum", "<u8")], align=True)
FR_DTYPE = np.dtype([("sse", "<u8", (3,)), ("mse", "<f8", (3,)), ("mse_avg", "<f8"), ("psnr", "<f8", (3,)),
                     ("psnr_avg", "<f8"), ("ssim", "<f8", (3,)), ("ssim_all", "<f8")], align=True)
assert FRAME_DTYPE.itemsize == 40 and FR_DTYPE.itemsize == 120

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libvqa_b200.so and declare every prototype of include/vqa_b200.h."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(SO_PATH):
            raise VqaError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). This package has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        vp, u8p, i32p, f64p = C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
        L.vqa_abi_version.restype = C.c_int
        L.vqa_init.argtypes = [C.c_int, C.POINTER(vp)]
        L.vqa_destroy.argtypes = [vp]
        L.vqa_destroy.restype = None
        L.vqa_last_error.argtypes = [vp]
        L.vqa_last_error.restype = C.c_char_p
        L.vqa_set_stream.argtypes = [vp, vp]
        L.vqa_sync.argtypes = [vp]
        L.vqa_kernel_launches.argtypes = [vp]
        L.vqa_kernel_launches.restype = C.c_uint64
        L.vqa_stage_ms.argtypes = [vp, C.c_char_p, f64p, C.POINTER(C.c_uint64)]
        L.vqa_reset_timers.argtypes = [vp, C.c_int]
        L.vqa_complexity_frames.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_size_t, u8p, C.c_int,
                                            C.POINTER(Cfg), vp]
        L.vqa_psnr_ssim_planar.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), i32p, i32p, i32p, C.c_int, C.c_int, vp]
        L.vqa_analyze_clip.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(Cfg), vp,
                                       C.POINTER(vp), C.POINTER(vp), i32p, i32p, i32p, C.c_int, vp]
        L.vqa_analyze_clip_yuv420.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), i32p, C.c_int, C.c_int, C.c_int,
                                              C.POINTER(vp), C.c_int, C.POINTER(Cfg), vp, vp]
        L.vqa_debug_yuv2bgr.argtypes = [vp, u8p, u8p, u8p, C.c_int, C.c_int, u8p]
        L.vqa_comm_unique_id.argtypes = [vp]
        L.vqa_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
        L.vqa_comm_destroy.argtypes = [vp]
        L.vqa_clip_reduce.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int]
        L.vqa_comm_halo_exchange.argtypes = [vp, vp, vp, vp, C.c_size_t]
        L.vqa_framerate_series.argtypes = [vp, vp, C.c_int, vp]
        L.vqa_ewm_partial.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_double, f64p]
        L.vqa_debug_gray.argtypes = [vp, u8p, C.c_int, C.c_int, u8p]
        L.vqa_debug_resize.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.vqa_debug_hist.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.vqa_debug_canny.argtypes = [vp, u8p, C.c_int, C.c_int, u8p]
        L.vqa_debug_orb.argtypes = [vp, u8p, C.c_int, C.c_int, vp]
        L.vqa_kernel_profile.argtypes = [vp, C.c_int]
        L.vqa_kernel_report.argtypes = [vp, C.c_char_p, C.c_size_t]
        L.vqa_debug_flow.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, vp]
        L.vqa_debug_dct.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, vp]
        L.vqa_orb_default_cfg.argtypes = [C.POINTER(OrbCfg)]
        L.vqa_orb_default_cfg.restype = None
        L.vqa_orb_describe.argtypes = [C.POINTER(OrbCfg), C.c_int, C.c_int, i32p, i32p, i32p]
        L.vqa_orb_detect.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(OrbCfg),
                                     vp, vp, vp, C.c_int]
        L.vqa_debug_orb_pyramid.argtypes = [vp, u8p, C.c_int, C.c_int, C.POINTER(OrbCfg), C.c_int, u8p]
        L.vqa_debug_exact_taps.argtypes = [C.c_int, C.c_int, vp]
        if L.vqa_abi_version() != ABI_VERSION:
            raise VqaError("libvqa_b200.so ABI version mismatch")
        _lib = L
        return L


def _is_torch_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _np_u8(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise TypeError(f"expected uint8 frames, got {a.dtype}")
    return np.ascontiguousarray(a)


class Context:
    """One per GPU (and per process).  Not thread-safe; guard with ``lock`` when shared."""

    def __init__(self, device: int | None = None):
        self.lib = load_library()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
            try:
                import torch
                if torch.cuda.is_available():
                    device = torch.cuda.current_device()
            except Exception:
                pass
        self.device = device
        h = C.c_void_p()
        rc = self.lib.vqa_init(device, C.byref(h))
        if rc != 0:
            raise VqaError(f"vqa_init({device}) failed ({rc}): {self.lib.vqa_last_error(None).decode()} "
                           "-- a B200 (sm_100) GPU is required; there is no CPU fallback")
        self.h = h
        self.lock = threading.RLock()

    def close(self):
        if getattr(self, "h", None):
            self.lib.vqa_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc, what):
        if rc != 0:
            raise VqaError(f"{what} failed ({rc}): {self.lib.vqa_last_error(self.h).decode()}")

    def use_torch_stream(self):
        import torch
        self._check(self.lib.vqa_set_stream(self.h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                    "vqa_set_stream")

    def kernel_launches(self) -> int:
        return int(self.lib.vqa_kernel_launches(self.h))

    def reset_timers(self, enable=True):
        self._check(self.lib.vqa_reset_timers(self.h, int(enable)), "vqa_reset_timers")

    def stage_ms(self, stage: str):
        ms, n = C.c_double(), C.c_uint64()
        self._check(self.lib.vqa_stage_ms(self.h, stage.encode(), C.byref(ms), C.byref(n)), "vqa_stage_ms")
        return ms.value, int(n.value)

    def sync(self):
        self._check(self.lib.vqa_sync(self.h), "vqa_sync")

    # ------------------------------------------------------------------ a1-a8
    def complexity_frames(self, frames, resize_width, resize_height, mask=M_ALL, halo=None, dct_impl=0, orb_size=None):
        """frames: (n,h,w,3) uint8 BGR -- numpy (host) or CUDA torch tensor.  Returns a structured
        array (FRAME_DTYPE) with one row per frame.  ``orb_size`` = (width, height) runs the full ORB
        pipeline on gray(resize(frame, orb_size)); None keeps the reference's hard-wired 64x64."""
        if _is_torch_tensor(frames):
            if not frames.is_cuda:
                frames = frames.numpy()
        if _is_torch_tensor(frames):
            import torch
            if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
                raise TypeError("expected a (n,h,w,3) uint8 CUDA tensor")
            if frames.device.index != self.device:
                raise TypeError("frames live on %s, this context drives cuda:%d" % (frames.device, self.device))
            frames = frames.contiguous()
            torch.cuda.current_stream(frames.device).synchronize()
            n, h, w, _ = frames.shape
            ptr, on_dev, stride = frames.data_ptr(), 1, h * w * 3
            hptr = None
            if halo is not None:
                if not (_is_torch_tensor(halo) and halo.is_cuda and halo.device == frames.device and halo.dtype == torch.uint8
                        and tuple(halo.shape) == (h, w, 3)):
                    raise TypeError("halo must be a (h,w,3) uint8 tensor on the frames' device")
                halo = halo.contiguous()
                hptr = halo.data_ptr()
            keep = (frames, halo)
        else:
            frames = _np_u8(frames)
            if frames.ndim == 3:
                frames = frames[None]
            if frames.ndim != 4 or frames.shape[-1] != 3:
                raise TypeError("expected (n,h,w,3) uint8 BGR frames")
            n, h, w, _ = frames.shape
            ptr, on_dev, stride = frames.ctypes.data, 0, h * w * 3
            hptr = None
            if halo is not None:
                halo = _np_u8(halo)
                if halo.shape != (h, w, 3):
                    raise TypeError("halo frame must match the frame size")
                hptr = halo.ctypes.data
            keep = (frames, halo)
        out = np.zeros(n, dtype=FRAME_DTYPE)
        ow, oh = (int(orb_size[0]), int(orb_size[1])) if orb_size else (0, 0)
        cfg = Cfg(int(resize_width), int(resize_height), int(mask), int(dct_impl), ow, oh)
        with self.lock:
            rc = self.lib.vqa_complexity_frames(self.h, C.c_void_p(ptr), n, h, w, stride,
                                                C.c_void_p(hptr) if hptr else None, on_dev, C.byref(cfg),
                                                C.c_void_p(out.ctypes.data))
        self._check(rc, "vqa_complexity_frames")
        del keep
        return out

    # ------------------------------------------------------------------ a8 at any size
    def orb_detect(self, gray, cfg: OrbCfg | None = None, keypoints: bool = False, kp_cap: int = 4096):
        """cv2.ORB_create(...).detectAndCompute(gray, None) up to the keypoint list.  gray: (h,w) or
        (n,h,w) uint8, numpy or CUDA tensor.  Returns (counts[n], level_counts[n,16]) and, with
        ``keypoints``, a list of KEYPOINT_DTYPE arrays (grouped by octave, unordered inside one)."""
        on_dev = int(_is_torch_tensor(gray) and gray.is_cuda)
        if on_dev:
            import torch
            g = gray.contiguous()
            if g.dim() == 2:
                g = g[None]
            if g.dtype != torch.uint8 or g.dim() != 3:
                raise TypeError("expected (n,h,w) uint8 gray frames")
            torch.cuda.current_stream(g.device).synchronize()
            ptr = g.data_ptr()
        else:
            g = _np_u8(gray)
            if g.ndim == 2:
                g = g[None]
            if g.ndim != 3:
                raise TypeError("expected (n,h,w) uint8 gray frames")
            ptr = g.ctypes.data
        n, h, w = (int(v) for v in g.shape)
        counts = np.zeros(n, np.int32)
        levels = np.zeros((n, ORB_MAX_LEVELS), np.int32)
        kps = np.zeros((n, kp_cap), KEYPOINT_DTYPE) if keypoints else None
        with self.lock:
            rc = self.lib.vqa_orb_detect(self.h, C.c_void_p(ptr), n, h, w, h * w, on_dev,
                                         C.byref(cfg) if cfg is not None else None, C.c_void_p(counts.ctypes.data),
                                         C.c_void_p(levels.ctypes.data),
                                         C.c_void_p(kps.ctypes.data) if keypoints else None, kp_cap if keypoints else 0)
        self._check(rc, "vqa_orb_detect")
        del g
        if keypoints:
            return counts, levels, [kps[i, :min(int(counts[i]), kp_cap)] for i in range(n)]
        return counts, levels

    def debug_orb_pyramid(self, gray, level, cfg: OrbCfg | None = None):
        g = _np_u8(gray)
        lw, lh, _ = orb_describe(g.shape[0], g.shape[1], cfg)
        out = np.empty((lh[level], lw[level]), np.uint8)
        self._check(self.lib.vqa_debug_orb_pyramid(self.h, g.ctypes.data, g.shape[0], g.shape[1],
                                                   C.byref(cfg) if cfg is not None else None, level, out.ctypes.data),
                    "vqa_debug_orb_pyramid")
        return out

    # ------------------------------------------------------------------ a13
    def psnr_ssim(self, main_planes, ref_planes):
        """main/ref: 3 planes each, [n,h_c,w_c] uint8 (numpy or CUDA tensors).  FR_DTYPE rows."""
        on_dev = int(_is_torch_tensor(main_planes[0]) and main_planes[0].is_cuda)
        keep, mp, rp, pw, ph, st = [], (C.c_void_p * 3)(), (C.c_void_p * 3)(), (C.c_int32 * 3)(), (C.c_int32 * 3)(), (C.c_int32 * 3)()
        n = None
        for i in range(3):
            a, b = main_planes[i], ref_planes[i]
            if on_dev:
                a, b = a.contiguous(), b.contiguous()
                shp = tuple(a.shape)
                mp[i], rp[i] = a.data_ptr(), b.data_ptr()
            else:
                a, b = _np_u8(a), _np_u8(b)
                shp = a.shape
                mp[i], rp[i] = a.ctypes.data, b.ctypes.data
            if len(shp) == 2:
                shp = (1,) + shp
            if tuple(b.shape)[-2:] != shp[-2:]:
                raise TypeError("main/ref plane shapes differ")
            n = shp[0] if n is None else n
            if shp[0] != n:
                raise TypeError("planes disagree on the frame count")
            ph[i], pw[i], st[i] = shp[1], shp[2], shp[2]
            keep += [a, b]
        if on_dev:
            import torch
            torch.cuda.current_stream().synchronize()
        out = np.zeros(n, dtype=FR_DTYPE)
        with self.lock:
            rc = self.lib.vqa_psnr_ssim_planar(self.h, mp, rp, pw, ph, st, n, on_dev, C.c_void_p(out.ctypes.data))
        self._check(rc, "vqa_psnr_ssim_planar")
        del keep
        return out

    def analyze_clip(self, frames, resize_width, resize_height, main_planes, ref_planes, mask=M_ALL, dct_impl=0,
                     orb_size=None):
        """Host-buffer fast path for one clip: complexity rows + PSNR/SSIM rows with one interleaved
        upload schedule (vqa_analyze_clip).  frames (n,h,w,3) uint8; planes 3 x [n_pairs,h_c,w_c] uint8."""
        frames = _np_u8(frames)
        if frames.ndim != 4 or frames.shape[-1] != 3:
            raise TypeError("expected (n,h,w,3) uint8 BGR frames")
        n, h, w, _ = frames.shape
        keep, mp, rp = [], (C.c_void_p * 3)(), (C.c_void_p * 3)()
        pw, ph, st = (C.c_int32 * 3)(), (C.c_int32 * 3)(), (C.c_int32 * 3)()
        npairs = None
        for i in range(3):
            a, b = _np_u8(main_planes[i]), _np_u8(ref_planes[i])
            if a.ndim != 3 or a.shape != b.shape:
                raise TypeError("planes must be [n_pairs,h_c,w_c] and main/ref must agree")
            npairs = a.shape[0] if npairs is None else npairs
            if a.shape[0] != npairs:
                raise TypeError("planes disagree on the pair count")
            mp[i], rp[i], ph[i], pw[i], st[i] = a.ctypes.data, b.ctypes.data, a.shape[1], a.shape[2], a.shape[2]
            keep += [a, b]
        rows = np.zeros(n, dtype=FRAME_DTYPE)
        fr = np.zeros(npairs, dtype=FR_DTYPE)
        ow, oh = (int(orb_size[0]), int(orb_size[1])) if orb_size else (0, 0)
        cfg = Cfg(int(resize_width), int(resize_height), int(mask), int(dct_impl), ow, oh)
        with self.lock:
            rc = self.lib.vqa_analyze_clip(self.h, C.c_void_p(frames.ctypes.data), n, h, w, h * w * 3, C.byref(cfg),
                                           C.c_void_p(rows.ctypes.data), mp, rp, pw, ph, st, npairs,
                                           C.c_void_p(fr.ctypes.data))
        self._check(rc, "vqa_analyze_clip")
        del keep
        return rows, fr

    def analyze_clip_yuv420(self, main_planes, ref_planes, resize_width, resize_height, mask=M_ALL, halo_planes=None,
                            dct_impl=0, orb_size=None):
        """Both halves of a clip from ONE set of yuv420p planes (vqa_analyze_clip_yuv420).
        main_planes = (Y [n,h,w], U [n,h/2,w/2], V [n,h/2,w/2]) of the ENCODED clip; the complexity metrics run
        on the BGR frames cv2.VideoCapture would decode from them (bit-exact libswscale conversion on the
        device).  ref_planes = the source clip's planes for PSNR/SSIM, or None.  numpy arrays (host) or CUDA
        tensors (all of them).  halo_planes = (Y [h,w], U, V) of the previous sampled frame, or None.
        Returns (rows, fr) -- fr is None without ref_planes."""
        on_dev = int(_is_torch_tensor(main_planes[0]) and main_planes[0].is_cuda)
        keep = []

        def ptrs(planes, frame_dims):
            arr, shapes = (C.c_void_p * 3)(), []
            for i in range(3):
                a = planes[i]
                if on_dev:
                    if not (_is_torch_tensor(a) and a.is_cuda and a.device.index == self.device):
                        raise TypeError("all planes must be CUDA tensors on the context's device")
                    a = a.contiguous()
                    if str(a.dtype) != "torch.uint8":
                        raise TypeError("planes must be uint8")
                    arr[i] = a.data_ptr()
                else:
                    a = _np_u8(a)
                    arr[i] = a.ctypes.data
                if len(a.shape) != frame_dims:
                    raise TypeError("planes must be [n,h_c,w_c] stacks (halo: [h_c,w_c])")
                shapes.append(tuple(int(v) for v in a.shape))
                keep.append(a)
            return arr, shapes

        mp, shp = ptrs(main_planes, 3)
        n, h, w = shp[0]
        if (h | w) & 1:
            raise VqaError("yuv420p frames must have even sizes")
        if shp[1] != (n, h // 2, w // 2) or shp[2] != (n, h // 2, w // 2):
            raise TypeError("U/V planes must be [n,h/2,w/2]")
        rp = None
        if ref_planes is not None:
            rp, rshp = ptrs(ref_planes, 3)
            if rshp != shp:
                raise TypeError("main/ref plane shapes differ")
        hp = None
        if halo_planes is not None:
            hp, hshp = ptrs(halo_planes, 2)
            if hshp != [(h, w), (h // 2, w // 2), (h // 2, w // 2)]:
                raise TypeError("halo planes must match the frame size")
        if on_dev:
            import torch
            torch.cuda.current_stream(self.device).synchronize()
        st = (C.c_int32 * 3)(w, w // 2, w // 2)
        rows = np.zeros(n, dtype=FRAME_DTYPE)
        fr = np.zeros(n, dtype=FR_DTYPE) if rp is not None else None
        ow, oh = (int(orb_size[0]), int(orb_size[1])) if orb_size else (0, 0)
        cfg = Cfg(int(resize_width), int(resize_height), int(mask), int(dct_impl), ow, oh)
        with self.lock:
            rc = self.lib.vqa_analyze_clip_yuv420(self.h, mp, rp, st, n, h, w, hp, on_dev, C.byref(cfg),
                                                  C.c_void_p(rows.ctypes.data),
                                                  C.c_void_p(fr.ctypes.data) if fr is not None else None)
        self._check(rc, "vqa_analyze_clip_yuv420")
        del keep
        return rows, fr

    # ------------------------------------------------------------------ e: multi-GPU close (NCCL inside the library)
    def comm_init(self, rank: int, world: int, id_bytes: bytes):
        """Create the context's own NCCL communicator (vqa_comm_init); ``id_bytes`` from comm_unique_id() of
        rank 0, distributed by the caller (e.g. torch.distributed.broadcast_object_list, a file, MPI)."""
        if len(id_bytes) != COMM_ID_BYTES:
            raise ValueError("NCCL unique id must be %d bytes" % COMM_ID_BYTES)
        buf = C.create_string_buffer(bytes(id_bytes), COMM_ID_BYTES)
        with self.lock:
            self._check(self.lib.vqa_comm_init(self.h, buf, int(rank), int(world)), "vqa_comm_init")
        self.comm_rank, self.comm_world = int(rank), int(world)

    def comm_destroy(self):
        with self.lock:
            self._check(self.lib.vqa_comm_destroy(self.h), "vqa_comm_destroy")
        self.comm_world = 0

    def clip_reduce(self, partials, ints):
        """Sum over all ranks (ONE ncclAllReduce on the context's stream): float64 partials and int64 totals."""
        p = np.ascontiguousarray(partials, dtype=np.float64).copy()
        i = np.ascontiguousarray(ints, dtype=np.int64).copy()
        with self.lock:
            rc = self.lib.vqa_clip_reduce(self.h, None, C.c_void_p(p.ctypes.data) if p.size else None, int(p.size),
                                          C.c_void_p(i.ctypes.data) if i.size else None, int(i.size))
        self._check(rc, "vqa_clip_reduce")
        return p.reshape(np.shape(partials)), i.reshape(np.shape(ints))

    def halo_exchange(self, send, recv):
        """send: CUDA uint8 tensor (this rank's last frame) or None; recv: CUDA uint8 tensor or None."""
        nbytes = int((send if send is not None else recv).numel())
        with self.lock:
            rc = self.lib.vqa_comm_halo_exchange(self.h, None, C.c_void_p(send.data_ptr()) if send is not None else None,
                                                 C.c_void_p(recv.data_ptr()) if recv is not None else None, nbytes)
        self._check(rc, "vqa_comm_halo_exchange")

    # ------------------------------------------------------------------ a9 / a10
    def framerate_series(self, timestamps_ms):
        ts = np.ascontiguousarray(timestamps_ms, dtype=np.float64)
        out = np.zeros(max(len(ts) - 1, 0), dtype=np.float64)
        with self.lock:
            rc = self.lib.vqa_framerate_series(self.h, C.c_void_p(ts.ctypes.data), len(ts), C.c_void_p(out.ctypes.data))
        self._check(rc, "vqa_framerate_series")
        return out

    def ewm_partial(self, x, offset=0, total=None, alpha=0.8) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        total = len(x) + offset if total is None else total
        out = C.c_double()
        with self.lock:
            rc = self.lib.vqa_ewm_partial(self.h, C.c_void_p(x.ctypes.data) if len(x) else None, len(x),
                                          int(offset), int(total), float(alpha), C.byref(out))
        self._check(rc, "vqa_ewm_partial")
        return out.value

    # ------------------------------------------------------------------ debug taps (tests)
    def debug_gray(self, frame):
        f = _np_u8(frame)
        out = np.empty(f.shape[:2], np.uint8)
        self._check(self.lib.vqa_debug_gray(self.h, f.ctypes.data, f.shape[0], f.shape[1], out.ctypes.data), "debug_gray")
        return out

    def debug_resize(self, img, rw, rh):
        s = _np_u8(img)
        cn = 1 if s.ndim == 2 else s.shape[2]
        out = np.empty((rh, rw) if s.ndim == 2 else (rh, rw, cn), np.uint8)
        self._check(self.lib.vqa_debug_resize(self.h, s.ctypes.data, s.shape[0], s.shape[1], cn, rw, rh, out.ctypes.data), "debug_resize")
        return out

    def debug_hist(self, frame, rw, rh):
        f = _np_u8(frame)
        out = np.zeros((4, 256), np.uint32)
        self._check(self.lib.vqa_debug_hist(self.h, f.ctypes.data, f.shape[0], f.shape[1], rw, rh, out.ctypes.data), "debug_hist")
        return out

    def debug_orb(self, frame):
        f = _np_u8(frame)
        out = np.zeros(117, np.int32)
        self._check(self.lib.vqa_debug_orb(self.h, f.ctypes.data, f.shape[0], f.shape[1], out.ctypes.data), "debug_orb")
        return out[:100].reshape(10, 10), out[100:116].reshape(4, 4), int(out[116])

    def kernel_profile(self, enable=True):
        self._check(self.lib.vqa_kernel_profile(self.h, int(enable)), "vqa_kernel_profile")

    def kernel_report(self):
        """{kernel: dict(launches, ms, bytes, flops)} since kernel_profile(True)."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self.lib.vqa_kernel_report(self.h, buf, len(buf)), "vqa_kernel_report")
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms, b, fl = line.rsplit(None, 4)
            name = name.replace(' ', '').strip('()')
            out[name] = dict(launches=int(n), ms=float(ms), bytes=float(b), flops=float(fl))
        return out

    def debug_yuv2bgr(self, y, u, v):
        y, u, v = _np_u8(y), _np_u8(u), _np_u8(v)
        h, w = y.shape
        if u.shape != (h // 2, w // 2) or v.shape != u.shape:
            raise TypeError("U/V planes must be [h/2,w/2]")
        out = np.empty((h, w, 3), np.uint8)
        self._check(self.lib.vqa_debug_yuv2bgr(self.h, y.ctypes.data, u.ctypes.data, v.ctypes.data, h, w, out.ctypes.data),
                    "debug_yuv2bgr")
        return out

    def debug_canny(self, gray):
        g = _np_u8(gray)
        out = np.empty_like(g)
        self._check(self.lib.vqa_debug_canny(self.h, g.ctypes.data, g.shape[0], g.shape[1], out.ctypes.data), "debug_canny")
        return out

    def debug_flow(self, prev_gray, next_gray):
        p, q = _np_u8(prev_gray), _np_u8(next_gray)
        out = np.empty(p.shape + (2,), np.float32)
        self._check(self.lib.vqa_debug_flow(self.h, p.ctypes.data, q.ctypes.data, p.shape[0], p.shape[1], out.ctypes.data), "debug_flow")
        return out

    def debug_dct(self, gray, impl=0):
        g = _np_u8(gray)
        out = np.empty(g.shape, np.float32)
        self._check(self.lib.vqa_debug_dct(self.h, g.ctypes.data, g.shape[0], g.shape[1], impl, out.ctypes.data), "debug_dct")
        return out


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it and distributes the 128 bytes)."""
    L = load_library()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = L.vqa_comm_unique_id(buf)
    if rc != 0:
        raise VqaError(f"vqa_comm_unique_id failed ({rc}): {L.vqa_last_error(None).decode()}")
    return buf.raw


def orb_describe(h: int, w: int, cfg: OrbCfg | None = None):
    """Host-only (no GPU): ORB pyramid level widths, heights and per-level feature quotas."""
    L = load_library()
    lw, lh, q = ((C.c_int32 * ORB_MAX_LEVELS)() for _ in range(3))
    n = L.vqa_orb_describe(C.byref(cfg) if cfg is not None else None, int(h), int(w), lw, lh, q)
    if n < 0:
        raise VqaError(f"vqa_orb_describe failed ({n})")
    return list(lw[:n]), list(lh[:n]), list(q[:n])


def exact_taps(src_len: int, dst_len: int):
    """Host-only: INTER_LINEAR_EXACT taps of the ORB pyramid as (offset, weight_right) arrays (8.8 fixed point)."""
    L = load_library()
    out = np.zeros(int(dst_len), np.uint32)
    rc = L.vqa_debug_exact_taps(int(src_len), int(dst_len), C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise VqaError(f"vqa_debug_exact_taps failed ({rc})")
    return (out >> 16).astype(np.int64), (out & 0xFFFF).astype(np.int64)


_contexts: dict = {}
_ctx_lock = threading.Lock()


def _resolve_device(device):
    if device is not None:
        return int(device)
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return int(os.environ.get("LOCAL_RANK", "0"))


def get_context(device: int | None = None, role: str = "complexity") -> Context:
    """Process-wide context cache keyed by (pid, device, role): one context (scratch arena, streams)
    per GPU and per half of the path.  The full-reference half (``role="fr"``) has its own context so
    a caller may run PSNR/SSIM of one clip in a second thread while the complexity pass of the same
    clip is in flight (its H2D copies then overlap the Farneback compute)."""
    key = (os.getpid(), _resolve_device(device), role)
    with _ctx_lock:
        ctx = _contexts.get(key)
        if ctx is None:
            ctx = Context(key[1])
            _contexts[key] = ctx
        return ctx
