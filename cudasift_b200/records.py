"""The 576-byte SiftPoint record (cudaSift.h:6-22) as a numpy dtype: pure numpy, no CUDA library involved."""
import numpy as np

# descriptor at byte 64
SIFT_DTYPE = np.dtype([
    ("xpos", "<f4"), ("ypos", "<f4"), ("scale", "<f4"), ("sharpness", "<f4"), ("edgeness", "<f4"),
    ("orientation", "<f4"), ("score", "<f4"), ("ambiguity", "<f4"), ("match", "<i4"),
    ("match_xpos", "<f4"), ("match_ypos", "<f4"), ("match_error", "<f4"), ("subsampling", "<f4"),
    ("empty", "<f4", (3,)), ("data", "<f4", (128,))])
assert SIFT_DTYPE.itemsize == 576
