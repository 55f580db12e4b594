"""Generate the golden fixtures under tests/golden/ by running the LIVE REFERENCE modules (imported from
/root/reference, which exists only in the build container) on oracle.weights state_dicts and oracle.inputs inputs.

    python tests/golden/make_golden.py

The fixtures pin the oracle restatement (oracle/ref_models.py) -- see tests/test_oracle_golden.py -- and, through
it, the CUDA path.  Shims needed to import the reference offline are the ones SURVEY.md App. B lists: the processor
and Wav2Vec2Model `from_pretrained` calls are replaced by default-config constructions, and eager attention is
selected because ref:src/model/wav2vec.py:101 forces output_attentions=True.
"""
from __future__ import annotations

import io
import os
import sys
import contextlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import inputs as oin            # noqa: E402
from oracle import weights as ow            # noqa: E402
from oracle import ref_models as orm        # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_grad_enabled(False)


def sub(t: torch.Tensor, step: int) -> np.ndarray:
    return t.reshape(-1)[::step].contiguous().numpy().astype(np.float32)


def build_reference_faceformer():
    from transformers import Wav2Vec2Config, Wav2Vec2FeatureExtractor
    import src.model.faceformer as ff
    from src.model.wav2vec import Wav2Vec2Model as RefW2V

    class _P:
        @staticmethod
        def from_pretrained(name):
            return Wav2Vec2FeatureExtractor()

    ff.Wav2Vec2Processor = _P

    def _mk(cls, name):
        cfg = Wav2Vec2Config()
        cfg._attn_implementation = "eager"
        return cls(cfg)

    RefW2V.from_pretrained = classmethod(_mk)
    return ff.Faceformer(15069, 12), ff


def check_keys(module: torch.nn.Module, sd, what: str):
    ref_sd = module.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()) or set(ref_sd.keys()) == set(sd.keys()), (
        what, set(ref_sd.keys()) ^ set(sd.keys()))
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(sd[k].shape), (what, k, tuple(v.shape), tuple(sd[k].shape))
        assert v.dtype == sd[k].dtype, (what, k, v.dtype, sd[k].dtype)
    module.load_state_dict(sd, strict=True)


def main():
    torch.manual_seed(0)
    meta = {}

    # ---------------- VOCA ----------------
    from src.model.voca import Voca
    m = Voca(15069, 12).eval()
    sd = ow.make_state_dict("voca", seed=11)
    check_keys(m, sd, "voca")
    B = 6
    x, oh, tp = oin.voca_features(B, 1), oin.one_hot(B, 12, 1), oin.batch_templates(B, 1)
    with contextlib.redirect_stdout(io.StringIO()):        # ref:voca.py:43 prints shapes
        y = m(x, oh, tp)
    y_or = orm.voca_forward(sd, x, oh, tp)
    print("voca   ref-vs-oracle max abs diff", float((y - y_or).abs().max()))
    np.savez_compressed(os.path.join(OUT, "voca.npz"), out=sub(y, 7), seed_w=11, seed_in=1, batch=B, step=7)

    # ---------------- Audio2Mesh (eval BN and train BN) ----------------
    from src.model.audio2face import Audio2Mesh
    m = Audio2Mesh(15069, 12).eval()
    sd = ow.make_state_dict("audio2mesh", seed=12)
    check_keys(m, sd, "audio2mesh")
    B = 4
    x, oh, tp = oin.a2m_features(B, 2), oin.one_hot(B, 12, 2), oin.batch_templates(B, 2)
    y_eval = m(x, oh, tp)
    m.train()
    y_train = m(x, oh, tp)
    m.eval()
    print("a2m    ref-vs-oracle eval ", float((y_eval - orm.audio2mesh_forward(sd, x, oh, tp)).abs().max()),
          " train-bn", float((y_train - orm.audio2mesh_forward(sd, x, oh, tp, train_bn=True)).abs().max()))
    np.savez_compressed(os.path.join(OUT, "audio2mesh.npz"), out_eval=sub(y_eval, 7), out_train=sub(y_train, 7),
                        seed_w=12, seed_in=2, batch=B, step=7)

    # ---------------- losses ----------------
    from src.loss import VocaLoss, FaceFormerLoss
    rows = 6
    tp = oin.batch_templates(rows, 3)
    pred = oin.gt_like((rows, 5023, 3), tp, 31)
    gt = oin.gt_like((rows, 5023, 3), tp, 32)
    lv = VocaLoss()(pred, gt)
    predf = oin.gt_like((1, 7, 5023, 3), tp[:1, None], 33)
    gtf = oin.gt_like((1, 7, 5023, 3), tp[:1, None], 34)
    lf = FaceFormerLoss()(predf, gtf)
    np.savez_compressed(os.path.join(OUT, "loss.npz"),
                        voca=np.array([float(lv["loss"]), float(lv["rec_loss"]), float(lv["vel_loss"])], dtype=np.float64),
                        faceformer=np.array([float(lf["loss"]), float(lf["rec_loss"]), float(lf["vel_loss"])], dtype=np.float64))
    lo = orm.voca_loss(pred, gt)
    print("loss   ref", float(lv["loss"]), "oracle", float(lo["loss"]))

    # ---------------- FaceFormer ----------------
    model, ff = build_reference_faceformer()
    model.eval()
    sd = ow.make_state_dict("faceformer", seed=13)
    check_keys(model, sd, "faceformer")
    # closed-form biased mask == reference construction
    bm_ref = ff.init_biased_mask(n_head=4, max_seq_len=600, period=60)
    bm_or = orm.init_biased_mask(4, 600, 60)
    assert torch.equal(bm_ref, bm_or), "biased mask closed form differs from the reference"
    assert torch.equal(model.PPE.pe, ow.ppe_table()), "PPE table differs from the reference"
    edm = ff.enc_dec_mask(torch.device("cpu"), "vocaset", 5, 9)
    assert torch.equal(edm, orm.enc_dec_mask(5, 9))

    fixtures = {}
    for tag, n_samples, seed_in in (("a", 16000, 5), ("b", 11200, 6)):
        audio = oin.audio(1, n_samples, seed_in)
        oh = oin.one_hot(1, 12, seed_in)
        tp = oin.batch_templates(1, seed_in, scale=100.0)        # centimetre convention of training_step
        y = model(audio, oh, tp)
        T = y.shape[1]
        hs = model.audio_encoder(orm.processor_normalize(audio[0])[None], "vocaset", frame_num=T).last_hidden_state
        y_or, parts = orm.faceformer_forward(sd, audio, oh, tp, return_parts=True)
        print(f"ff[{tag}] T={T} ref-vs-oracle out", float((y - y_or).abs().max()), " encoder",
              float((hs - parts["encoder"]).abs().max()), " |offset|max", float((y - tp[:, None]).abs().max()))
        fixtures[f"out_{tag}"] = sub(y, 31)
        fixtures[f"enc_{tag}"] = sub(hs, 5)
        fixtures[f"n_{tag}"] = n_samples
        fixtures[f"seed_{tag}"] = seed_in
    np.savez_compressed(os.path.join(OUT, "faceformer.npz"), seed_w=13, step_out=31, step_enc=5, **fixtures)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
