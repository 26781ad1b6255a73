"""bench.py control flow without a GPU: every device call is replaced by a stub, so that the step accounting of both
arms (passes per step, images counted, identical `config` objects, bytes per step) is checked on the CPU."""
import contextlib
import ctypes
import importlib.util
import io
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeFn:
    def __init__(self, name, lib):
        self.name, self.lib, self.restype, self.argtypes = name, lib, None, None

    def __call__(self, *a):
        self.lib.calls[self.name] = self.lib.calls.get(self.name, 0) + 1
        if "ExtractSift" in self.name:
            a[0]._obj.numPts = 1710
        if "MatchSiftData" in self.name:
            return 0.2
        if "AllocSiftTempMemory" in self.name:
            return 1234
        return 0


class _FakeLib:
    def __init__(self):
        self.calls, self.fns = {}, {}

    def __getattr__(self, n):
        if n.startswith("_Z") or n.startswith("cuda"):
            return self.__dict__["fns"].setdefault(n, _FakeFn(n, self))
        raise AttributeError(n)


def _load_bench(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    small = np.zeros((bench.H, bench.W), np.float32)
    monkeypatch.setattr(bench, "bind_numa", lambda i: {"bound": False})
    monkeypatch.setattr(bench, "make_images", lambda rank, distinct, standalone=False: [small] * distinct)
    return bench


def _args(impl):
    return types.SimpleNamespace(gpus=1, steps=20, warmup=5, impl=impl, batch=32, rounds=0, streams=2, distinct=8, no_cpu=True)


def test_both_arms_count_the_same_work(monkeypatch):
    bench = _load_bench(monkeypatch)
    # ---- reference arm: the library handle is a stub
    fake = _FakeLib()
    monkeypatch.setattr(bench.ctypes, "CDLL", lambda *a, **k: fake)
    monkeypatch.setattr(bench, "load_cudart", lambda: _FakeLib())

    class SM:
        @staticmethod
        def synth_descriptors(n, seed):
            return np.zeros(n, np.dtype([("x", "f4", 144)]))
    monkeypatch.setattr(bench, "synth_module", lambda standalone=False: SM)
    monkeypatch.setattr(bench.os.path, "exists", lambda p: True)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.run_reference(_args("reference"))
    ref = json.loads(buf.getvalue().strip().splitlines()[-1])
    extracts = [v for k, v in fake.calls.items() if "ExtractSift" in k][0]
    assert ref["config"]["images_per_step_per_gpu"] == 384                  # 12 passes over 32 images
    assert extracts == (5 + 20 + 10) * 384                                  # warm-up + timed + e2e steps
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 384 * bench.W * bench.H * 4

    # ---- product arm: the package is a stub
    class L:
        n = 0
        buf = (ctypes.c_float * (bench.W * bench.H))()
        def cs_device_sync(self): pass
        def cs_max_batch(self): return 32
        def cs_event_create(self): return 1
        def cs_event_record(self, e, h): pass
        def cs_event_elapsed_ms(self, a, b): return 320.0
        def cs_launch_count(self): return L.n
        def cs_extractor_host_image_at(self, h, i): return ctypes.addressof(L.buf)
    lib = L()

    class Img:
        d_data = 1
        def Allocate(self, *a): return self
        def Download(self): return 0.0

    class Ex:
        submitted = 0
        def __init__(self, *a, **k): self.handle = 1
        def submit_device_batch(self, ptrs, *a):
            L.n += 6
            Ex.submitted += len(ptrs)
        def submit_host_batch(self, ptrs, *a): pass
        def wait_batch(self, b): return [1710] * b
    cs = types.SimpleNamespace(InitCuda=lambda d: None, lib=lambda: lib, CudaImage=Img, Extractor=Ex,
                               iAlignUp=lambda a, b: a if a % b == 0 else a - a % b + b)
    monkeypatch.setitem(sys.modules, "cudasift_b200", cs)
    for f in ("bench_dropin", "bench_roofline", "bench_match"):
        monkeypatch.setattr(bench, f, lambda *a, **k: {})
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.run_product(_args("b200"))
    prod = json.loads(buf.getvalue().strip().splitlines()[-1])
    assert Ex.submitted == (5 + 20) * 384
    assert abs(prod["value"] - 20 * 384 / 0.320) < 1.0                       # images of the timed steps / device time
    assert prod["config"] == ref["config"]                                   # what the driver compares (same_config)
    assert prod["e2e"]["images"] >= 512 and prod["e2e"]["h2d_bytes_per_step"] == 384 * bench.W * bench.H * 4
    assert prod["gpu_launches"] == 20 * 12 * 2 * 6
