"""ctypes driver for a library exporting the reference's C++ API (Itanium-mangled symbols).

Works for oracle/_ref/libcudasift_ref.so (the unmodified reference, when it travelled to the
GPU box) and for libcudasift_b200.so itself -- both export the same symbols, so the same
harness feeds both with identical inputs.  References are passed as
pointers; structs are laid out as in cudaSift.h / cudaImage.h.
"""
import ctypes
import os

import numpy as np

from cudasift_b200 import SIFT_DTYPE
from cudasift_b200 import build as _build

REF_LIB = _build.REF_LIB


class CSiftData(ctypes.Structure):
    _fields_ = [("numPts", ctypes.c_int), ("maxPts", ctypes.c_int), ("h_data", ctypes.c_void_p),
                ("d_data", ctypes.c_void_p)]


class CCudaImage(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("pitch", ctypes.c_int),
                ("h_data", ctypes.c_void_p), ("d_data", ctypes.c_void_p), ("t_data", ctypes.c_void_p),
                ("d_internalAlloc", ctypes.c_bool), ("h_internalAlloc", ctypes.c_bool)]


assert ctypes.sizeof(CSiftData) == 24 and ctypes.sizeof(CCudaImage) == 48


class quiet_stdout:
    """The reference printf()s on every call (quirk Q13): silence fd 1 meanwhile."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        import ctypes as c
        try:
            c.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


class CxxSiftLib:
    def __init__(self, path, device=0):
        self.path = path
        L = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        self.L = L
        c = ctypes
        P = c.POINTER
        self.InitCuda = L._Z8InitCudai
        self.InitCuda.argtypes = [c.c_int]
        self._alloc = L._Z19AllocSiftTempMemoryiiib
        self._alloc.restype, self._alloc.argtypes = c.c_void_p, [c.c_int, c.c_int, c.c_int, c.c_bool]
        self._free = L._Z18FreeSiftTempMemoryPf
        self._free.argtypes = [c.c_void_p]
        self._extract = L._Z11ExtractSiftR8SiftDataR9CudaImageidffbPf
        self._extract.argtypes = [P(CSiftData), P(CCudaImage), c.c_int, c.c_double, c.c_float, c.c_float, c.c_bool,
                                  c.c_void_p]
        self._init = L._Z12InitSiftDataR8SiftDataibb
        self._init.argtypes = [P(CSiftData), c.c_int, c.c_bool, c.c_bool]
        self._freedata = L._Z12FreeSiftDataR8SiftData
        self._freedata.argtypes = [P(CSiftData)]
        self._match = L._Z13MatchSiftDataR8SiftDataS0_
        self._match.restype, self._match.argtypes = c.c_double, [P(CSiftData), P(CSiftData)]
        self._imgalloc = L._ZN9CudaImage8AllocateEiiibPfS0_
        self._imgalloc.argtypes = [P(CCudaImage), c.c_int, c.c_int, c.c_int, c.c_bool, c.c_void_p, c.c_void_p]
        self._imgdown = L._ZN9CudaImage8DownloadEv
        self._imgdown.restype, self._imgdown.argtypes = c.c_double, [P(CCudaImage)]
        self._imgread = L._ZN9CudaImage8ReadbackEv
        self._imgread.restype, self._imgread.argtypes = c.c_double, [P(CCudaImage)]
        self._imgdtor = L._ZN9CudaImageD1Ev
        self._imgdtor.argtypes = [P(CCudaImage)]
        # stage-level host functions (cudaSiftH.h:11-22) -- present in both libraries
        self._lowpass = L._Z7LowPassR9CudaImageS0_f
        self._lowpass.restype, self._lowpass.argtypes = c.c_double, [P(CCudaImage), P(CCudaImage), c.c_float]
        self._scaledown = L._Z9ScaleDownR9CudaImageS0_f
        self._scaledown.restype, self._scaledown.argtypes = c.c_double, [P(CCudaImage), P(CCudaImage), c.c_float]
        self._scaleup = L._Z7ScaleUpR9CudaImageS0_
        self._scaleup.restype, self._scaleup.argtypes = c.c_double, [P(CCudaImage), P(CCudaImage)]
        try:   # internal stage launcher of the reference (cudaSiftH.h:20); absent from the product
            self._laplace = L._Z12LaplaceMultiyR9CudaImagePS_i
            self._laplace.restype = c.c_double
            self._laplace.argtypes = [c.c_ulonglong, P(CCudaImage), P(CCudaImage), c.c_int]
        except AttributeError:
            self._laplace = None
        with quiet_stdout():
            self.InitCuda(device)     # one process per GPU: the reference keeps per-process state (quirk Q12)

    # ---- images ----
    def image(self, arr):
        """Device image from a host float array (pitch = iAlignUp(w,128) as mainSift.cpp:51)."""
        arr = np.ascontiguousarray(arr, np.float32)
        h, w = arr.shape
        img = CCudaImage()
        p = w if w % 128 == 0 else w - w % 128 + 128
        self._imgalloc(ctypes.byref(img), w, h, p, False, None, arr.ctypes.data_as(ctypes.c_void_p))
        img._keep = arr
        self._imgdown(ctypes.byref(img))
        return img

    def blank(self, w, h):
        img = CCudaImage()
        p = w if w % 128 == 0 else w - w % 128 + 128
        host = np.zeros((h, w), np.float32)
        self._imgalloc(ctypes.byref(img), w, h, p, False, None, host.ctypes.data_as(ctypes.c_void_p))
        img._keep = host
        return img

    def readback(self, img):
        self._imgread(ctypes.byref(img))
        return img._keep.copy()

    def free_image(self, img):
        self._imgdtor(ctypes.byref(img))

    # ---- stages ----
    def lowpass(self, arr, sigma):
        src, dst = self.image(arr), self.blank(arr.shape[1], arr.shape[0])
        with quiet_stdout():
            self._lowpass(ctypes.byref(dst), ctypes.byref(src), sigma)
        out = self.readback(dst)
        self.free_image(src); self.free_image(dst)
        return out

    def scaledown(self, arr):
        h, w = arr.shape
        src, dst = self.image(arr), self.blank(w // 2, h // 2)
        with quiet_stdout():
            self._scaledown(ctypes.byref(dst), ctypes.byref(src), 0.5)
        out = self.readback(dst)
        self.free_image(src); self.free_image(dst)
        return out

    def scaleup(self, arr):
        h, w = arr.shape
        src, dst = self.image(arr), self.blank(2 * w, 2 * h)
        with quiet_stdout():
            self._scaleup(ctypes.byref(dst), ctypes.byref(src))
        out = self.readback(dst)
        self.free_image(src); self.free_image(dst)
        return out

    def dog(self, arr, numOctaves, octave):
        """7 DoG planes of the reference's LaplaceMulti for `arr` taken as an octave base image.
        The taps live in a __constant__ array that only ExtractSift uploads (cudaSiftH.cu:111),
        so a throw-away extraction with the same numOctaves runs first."""
        import cudasift_b200 as cs
        assert self._laplace is not None
        h, w = arr.shape
        self.extract(np.ascontiguousarray(arr[:64, :128]) if h >= 64 and w >= 128 else arr, numOctaves=numOctaves)
        base = self.image(arr)
        p = base.pitch
        buf = cs.DeviceBuffer(7 * h * p * 4)
        res = (CCudaImage * 8)()
        for i in range(8):
            res[i].width, res[i].height, res[i].pitch = w, h, p
            res[i].d_data = buf.ptr + min(i, 6) * h * p * 4
        with quiet_stdout():
            self._laplace(0, ctypes.byref(base), res, octave)
        cs.lib().cs_device_sync()
        out = buf.download(np.float32, 7 * h * p).reshape(7, h, p)[:, :, :w].copy()
        buf.free(); self.free_image(base)
        return out

    # ---- extraction / matching ----
    def extract(self, arr, numOctaves=5, initBlur=1.0, thresh=3.0, lowestScale=0.0, scaleUp=False, maxPts=32768,
                use_temp=True, repeat=1):
        h, w = arr.shape
        img = self.image(arr)
        sd = CSiftData()
        self._init(ctypes.byref(sd), maxPts, True, True)
        tmp = self._alloc(w, h, numOctaves, scaleUp) if use_temp else None
        with quiet_stdout():
            for _ in range(repeat):
                self._extract(ctypes.byref(sd), ctypes.byref(img), numOctaves, initBlur, thresh, lowestScale, scaleUp,
                              tmp)
        n = sd.numPts
        buf = (ctypes.c_char * (n * SIFT_DTYPE.itemsize)).from_address(sd.h_data) if n else b""
        pts = np.frombuffer(buf, dtype=SIFT_DTYPE, count=n).copy() if n else np.zeros(0, SIFT_DTYPE)
        if tmp:
            self._free(tmp)
        self._freedata(ctypes.byref(sd))
        self.free_image(img)
        return pts

    def match(self, s1, s2):
        """MatchSiftData on host record arrays; returns the updated copy of s1 and the ms."""
        import cudasift_b200 as cs
        s1 = np.ascontiguousarray(s1, SIFT_DTYPE).copy()
        s2 = np.ascontiguousarray(s2, SIFT_DTYPE)
        d1, d2 = CSiftData(), CSiftData()
        self._init(ctypes.byref(d1), max(len(s1), 1) + 64, True, True)   # +64: the reference writes past n1 (Q8)
        self._init(ctypes.byref(d2), max(len(s2), 1) + 64, False, True)
        d1.numPts, d2.numPts = len(s1), len(s2)
        cs.lib().cs_memcpy_h2d(d1.d_data, s1.ctypes.data_as(ctypes.c_void_p), s1.nbytes)
        cs.lib().cs_memcpy_h2d(d2.d_data, s2.ctypes.data_as(ctypes.c_void_p), s2.nbytes)
        with quiet_stdout():
            ms = self._match(ctypes.byref(d1), ctypes.byref(d2))
        out = np.zeros(len(s1), SIFT_DTYPE)
        cs.lib().cs_memcpy_d2h(out.ctypes.data_as(ctypes.c_void_p), d1.d_data, out.nbytes)
        self._freedata(ctypes.byref(d1)); self._freedata(ctypes.byref(d2))
        return out, ms


def load_reference():
    """The unmodified reference library, or None when it did not travel / was not built."""
    if os.path.exists(REF_LIB):
        try:
            return CxxSiftLib(REF_LIB)
        except OSError:
            return None
    return None
