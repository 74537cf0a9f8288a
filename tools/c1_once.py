import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, ainmf
g = np.load("tests/golden/c1_part0.npz")
lab = ainmf.SpectralInpainter.__new__(ainmf.SpectralInpainter)
ainmf.SpectralInpainter.__init__(lab, filename=None, duration=0.05)
lab.sr = int(g["sr"]); lab.raw_audio = g["raw"].copy(); lab.apply_mask(0.2)
lab.restore_with_nmf(n_components=40, n_iter=50)
torch.cuda.synchronize()
print("ok", lab.n_iter_)
