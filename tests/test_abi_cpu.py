"""CPU checks of the drop-in boundary: the library loads without a GPU, exports every symbol
include/cudasift_b200.h declares and the reference's C++ (Itanium-mangled) API, reports
the missing device loudly (no CPU fallback), and keeps the ABI struct layouts."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import cudasift_b200 as cs
from cudasift_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the reference's public + stage-level symbols (nm on its objects)
MANGLED = """_Z8InitCudai _Z19AllocSiftTempMemoryiiib _Z18FreeSiftTempMemoryPf
_Z11ExtractSiftR8SiftDataR9CudaImageidffbPf _Z12InitSiftDataR8SiftDataibb _Z12FreeSiftDataR8SiftData
_Z13PrintSiftDataR8SiftData _Z13MatchSiftDataR8SiftDataS0_ _Z14FindHomographyR8SiftDataPfPiifff
_ZN9CudaImageC1Ev _ZN9CudaImageC2Ev _ZN9CudaImageD1Ev _ZN9CudaImageD2Ev _ZN9CudaImage8AllocateEiiibPfS0_
_ZN9CudaImage8DownloadEv _ZN9CudaImage8ReadbackEv _ZN9CudaImage11InitTextureEv
_ZN9CudaImage13CopyToTextureERS_b _Z6iDivUpii _Z8iDivDownii _Z8iAlignUpii _Z10iAlignDownii
_Z7ScaleUpR9CudaImageS0_ _Z9ScaleDownR9CudaImageS0_f _Z7LowPassR9CudaImageS0_f
_Z21PrepareLaplaceKernelsifPf""".split()


def _exports():
    out = subprocess.run(["nm", "-D", "--defined-only", build.build_library()], stdout=subprocess.PIPE, text=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_library_exports_c_abi():
    header = open(os.path.join(ROOT, "include", "cudasift_b200.h")).read()
    declared = set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    missing = declared - _exports()
    assert not missing, "declared in cudasift_b200.h but not exported: %s" % sorted(missing)


def test_library_exports_reference_cxx_api():
    missing = set(MANGLED) - _exports()
    assert not missing, "reference API symbols not exported: %s" % sorted(missing)


def test_managed_flavour_exports_the_same_api():
    """cudaSift.h:27-32 (MANAGEDMEM): the unified-memory flavour is a second library with the same mangled symbols."""
    lib = build.build_library(managed=True)
    out = subprocess.run(["nm", "-D", "--defined-only", lib], stdout=subprocess.PIPE, text=True).stdout
    syms = {line.split()[-1] for line in out.splitlines() if line.strip()}
    assert not (set(MANGLED) - syms)
    assert os.path.exists(build.build_demo(managed=True))


def test_struct_layouts():
    assert cs.SIFT_DTYPE.itemsize == 576
    assert cs.SIFT_DTYPE.fields["data"][1] == 64 and cs.SIFT_DTYPE.fields["match"][1] == 32
    assert cs.SIFT_DTYPE.fields["subsampling"][1] == 48
    import reflib
    assert ctypes.sizeof(reflib.CSiftData) == 24 and ctypes.sizeof(reflib.CCudaImage) == 48


def test_no_cpu_fallback():
    """Without a GPU every entry point must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = cs.lib()
    assert L.cs_init(0) == -4                                   # CS_E_NODEV
    assert b"no CPU fallback" in L.cs_last_error()
    with pytest.raises(cs.CudaSiftError):
        cs.extract_host(np.zeros((64, 64), np.float32))
    with pytest.raises(cs.CudaSiftError):
        cs.InitCuda(0)


def test_int_helpers_and_temp_size():
    L = cs.lib()
    f = L._Z8iAlignUpii; f.argtypes = [ctypes.c_int, ctypes.c_int]
    assert f(1920, 128) == 1920 and f(1921, 128) == 2048 and cs.iAlignUp(500, 128) == 512
    g = L._Z6iDivUpii; g.argtypes = [ctypes.c_int, ctypes.c_int]
    assert g(1080, 32) == 34
    # arena sizes quoted in SURVEY.md 3.2: 60.1 MB (1280x960), 101.2 MB (1920x1080)
    assert abs(L.cs_temp_floats(1280, 960, 5, 0) * 4 / 1e6 - 60.1) < 0.3
    assert abs(L.cs_temp_floats(1920, 1080, 5, 0) * 4 / 1e6 - 101.2) < 0.3
    # level-0/1 kernel, ScaleDown chain, detector, cap fix-up, describe: 5 launches for any batch (round 1: 7 per image)
    assert L.cs_extract_launches_per_image(5, 0) == 5


def test_product_does_not_touch_the_oracle():
    """The product path must never import, link or call anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cudasift_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and f != "build.py":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
                assert not re.search(r"#\s*include\s*[<\"][^>\"]*oracle", text), f
                assert "liboracle" not in text and "dlopen" not in text, f
    needed = subprocess.run(["ldd", build.LIB], stdout=subprocess.PIPE, text=True).stdout
    assert "liboracle" not in needed
