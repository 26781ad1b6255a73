"""Debug helper (GPU box): reference vs product ExtractSift records on one synthetic image."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cudasift_b200 as cs, reflib
from cudasift_b200.synth import synth_image
cs.InitCuda(0)
ref = reflib.load_reference()
img = synth_image(1280, 960, seed=1000)
r = ref.extract(img, thresh=3.0)
m = cs.extract_host(img, thresh=3.0)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dump_extract.npz"), ref=r, mine=m)
print("dumped", len(r), len(m))
