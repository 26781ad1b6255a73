import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudasift_b200 as cs
from cudasift_b200.synth import synth_image
cs.InitCuda(0)
w, h = int(sys.argv[1]), int(sys.argv[2])
img = synth_image(w, h, seed=7)
a = cs.extract_host(img)
print("new:", len(a))
cs.set_tuning("legacy", 1)
b = cs.extract_host(img)
print("legacy:", len(b))
