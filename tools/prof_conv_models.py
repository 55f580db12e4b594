"""One bf16-precision forward of Audio2Mesh (64 windows) and Song2Face (64 windows) from raw audio windows (MFCC extractor
in front), a few times: target for ncu launch lists of the conv-model paths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from a2f_b200 import features, modules
from oracle import inputs as oin, weights as ow

dev = torch.device("cuda:0")
B = 64
ext = features.MFCCExtractor(22000, 32, 52, 440, None, 1024).to(dev).set_precision("bf16")
a2m = modules.Audio2Mesh(15069, 12); a2m.load_state_dict(ow.make_state_dict("audio2mesh", 12), strict=True)
s2f = modules.Song2Face(15069, 12); s2f.load_state_dict(ow.make_state_dict("song2face", 14), strict=True)
a2m, s2f = a2m.to(dev).eval().set_precision("bf16"), s2f.to(dev).eval().set_precision("bf16")
x, oh, tp = oin.speech_like_windows(B, seed=1).to(dev), oin.one_hot(B, 12, 1).to(dev), oin.batch_templates(B, 1).to(dev)
with torch.no_grad():
    for i in range(3):
        if i == 2:
            torch.cuda.synchronize(); torch.cuda.profiler.start()
        f = ext(x)
        a2m(f, oh, tp)
        s2f(f, oh, tp)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
