import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat.frontend import FbankFrontend
wave = (torch.randn(64, 163840) * 0.1).cuda()
fe = FbankFrontend(target_length=1024)
for _ in range(3): out = fe(wave)
torch.cuda.synchronize(); print("ok")
