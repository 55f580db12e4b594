"""B200-native (sm_100a) audio->mesh hot path of xtliu97/audio2face-pytorch.

Import as `a2f_b200` (the top-level shim package; this directory's name is not a Python identifier).
Contents: csrc/ (CUDA kernels + C-ABI), build.py (nvcc driver), lib.py (ctypes binding), ops.py (tensor wrappers),
modules.py (drop-in nn.Modules: Voca, Audio2Mesh, Faceformer, VocaLoss, FaceFormerLoss).
"""
from . import lib  # noqa: F401
from .lib import A2FError  # noqa: F401


def get_model(modelname: str):
    """Mirror of ref:src/model/lightning_model.py:50-58 for the models on the hot path."""
    from . import modules
    table = {"voca": modules.Voca}
    for opt in ("Audio2Mesh", "Faceformer", "Song2Face"):
        if hasattr(modules, opt):
            table[{"Audio2Mesh": "audio2mesh", "Faceformer": "faceformer", "Song2Face": "song2face"}[opt]] = getattr(modules, opt)
    if modelname not in table:
        raise KeyError(f"model {modelname!r} is outside the B200 hot path (SURVEY.md section 8)")
    return table[modelname]


def get_loss_fn(modelname: str):
    """Mirror of ref:src/model/lightning_model.py:70-73."""
    from . import modules
    return modules.FaceFormerLoss() if modelname == "faceformer" else modules.VocaLoss()


def get_extractor(extractor):
    """Mirror of ref:src/model/lightning_model.py:61-67 for the extractors on the B200 path ("mfcc", "wav2vec"; None -> no extractor)."""
    from . import features
    if extractor is None:
        return lambda *args, **kwargs: None
    table = {"mfcc": features.MFCCExtractor, "wav2vec": features.Wav2VecExtractor}
    if extractor not in table:
        raise KeyError(f"extractor {extractor!r} is outside the B200 hot path (SURVEY.md section 8(f))")
    return table[extractor]
