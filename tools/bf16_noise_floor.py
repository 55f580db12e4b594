"""CPU experiment (no CUDA code involved): the bf16 noise floor of FaceFormer at random init.

Emulates bf16 rounding of GEMM operands / stored activations INSIDE THE ORACLE's wav2vec2 encoder (fp32 accumulation, fp32
LayerNorm / softmax statistics, fp32 decoder -- the arithmetic contract of the CUDA bf16 path and of torch.autocast) and
measures, for one 5 s utterance of the bench batch, the vertex error and the relative error of FaceFormerLoss and its two
components against the fp32 oracle.  Modes:
  cur      the round-1 CUDA path: pre-LayerNorm sums AND LayerNorm outputs stored in bf16
  fusedln  LayerNorm applied to the un-rounded fp32 sum (what a GEMM epilogue with a fused LayerNorm does); bf16 outputs
  fp32res  additionally an fp32 residual stream (torch.autocast semantics)
Result (profiles/r2_bf16_noise_floor.txt): the 12 encoder layers carry ~95 % of the error; not rounding the pre-LN sum takes
the vertex error from 1.6 % to 1.0 % of the largest offset at T=300; rec_loss stays 2e-4 .. 4e-4 off in every mode -- that is
the floor of bf16 operands, not something a kernel can remove.     python tools/bf16_noise_floor.py [fps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import inputs as oin, ref_models as orm, weights as ow

r = lambda t: t.bfloat16().float()
def lin(x, w, b=None):
    y = r(x) @ r(w).t()
    return y if b is None else y + b

def gelu_tanh(x):
    return 0.5 * x * (1 + torch.tanh(x * (0.7978845608 + 0.0356774081 * x * x)))

def encoder_sim(sd, audio_norm, T, mode, gelu=F.gelu):
    p = "audio_encoder.feature_extractor.conv_layers."
    h = F.conv1d(audio_norm[:, None], sd[p + "0.conv.weight"], None, stride=5)
    h = r(gelu(F.group_norm(h, 512, sd[p + "0.layer_norm.weight"], sd[p + "0.layer_norm.bias"], eps=1e-5)))
    for i in range(1, 7):
        h = r(gelu(F.conv1d(h, r(sd[p + f"{i}.conv.weight"]), None, stride=2)))
    h = orm.linear_interpolation(h.transpose(1, 2), T)
    fp = "audio_encoder.feature_projection."
    h = r(F.layer_norm(h, (512,), sd[fp + "layer_norm.weight"], sd[fp + "layer_norm.bias"], 1e-5))
    h0 = lin(h, sd[fp + "projection.weight"], sd[fp + "projection.bias"])
    e = "audio_encoder.encoder."
    if mode == "cur":
        h0 = r(h0)
    pos = F.conv1d(r(h0).transpose(1, 2), r(orm.pos_conv_weight(sd)), sd[e + "pos_conv_embed.conv.bias"], padding=64, groups=16)
    pre = h0 + gelu(pos[:, :, :-1]).transpose(1, 2)
    if mode in ("cur",):
        pre = r(pre)
    h32 = F.layer_norm(pre, (768,), sd[e + "layer_norm.weight"], sd[e + "layer_norm.bias"], 1e-5)
    B, T_, _ = h32.shape
    for l in range(12):
        q_ = e + f"layers.{l}."
        hres = h32 if mode == "fp32res" else r(h32)
        qkv = [r(lin(h32, sd[q_ + f"attention.{n}_proj.weight"], sd[q_ + f"attention.{n}_proj.bias"])).view(B, T_, 12, 64).transpose(1, 2) for n in "qkv"]
        w = torch.softmax(qkv[0] @ qkv[1].transpose(2, 3) * 0.125, -1)
        a = r((r(w) @ qkv[2]).transpose(1, 2).reshape(B, T_, 768))
        pre1 = lin(a, sd[q_ + "attention.out_proj.weight"], sd[q_ + "attention.out_proj.bias"]) + hres
        if mode == "cur": pre1 = r(pre1)
        h1 = F.layer_norm(pre1, (768,), sd[q_ + "layer_norm.weight"], sd[q_ + "layer_norm.bias"], 1e-5)
        h1res = h1 if mode == "fp32res" else r(h1)
        f = r(gelu(lin(h1, sd[q_ + "feed_forward.intermediate_dense.weight"], sd[q_ + "feed_forward.intermediate_dense.bias"])))
        pre2 = lin(f, sd[q_ + "feed_forward.output_dense.weight"], sd[q_ + "feed_forward.output_dense.bias"]) + h1res
        if mode == "cur": pre2 = r(pre2)
        h32 = F.layer_norm(pre2, (768,), sd[q_ + "final_layer_norm.weight"], sd[q_ + "final_layer_norm.bias"], 1e-5)
    return r(h32)


def full(sd, audio, oh, tp, fps, mode, gelu=F.gelu):
    T = audio.shape[1] * fps // 16000
    an = orm.processor_normalize(audio.squeeze(0))[None]
    hs = orm.audio_encoder(sd, an, T) if mode == "exact" else encoder_sim(sd, an, T, mode, gelu)
    if mode == "exact":
        memory = F.linear(hs, sd["audio_feature_map.weight"], sd["audio_feature_map.bias"])
    else:
        memory = lin(hs, sd["audio_feature_map.weight"], sd["audio_feature_map.bias"])
    vo = orm.faceformer_decode(sd, memory, oh, T)
    return (vo + tp.reshape(1, 1, -1)).view(1, T, -1, 3), hs

if __name__ == "__main__":
    torch.set_grad_enabled(False)
    sd = ow.make_state_dict("faceformer", 13)
    fps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    n = 80000
    T = n * fps // 16000
    tpb = oin.batch_templates(8, 100, scale=100.0)
    audio, oh, tp = oin.audio(8, n, 100)[:1], oin.one_hot(8, 12, 100)[:1], tpb[:1]
    gt = oin.gt_like((8, T, 5023, 3), tpb[:, None], 200, scale=100.0)[:1]
    want, hs0 = full(sd, audio, oh, tp, fps, "exact")
    l0 = orm.faceformer_loss(want, gt)
    off = float((want - tp[:, None]).abs().max())
    print(f"T={T} |offset|max {off:.3f} cm; oracle loss {float(l0['loss']):.6f} rec {float(l0['rec_loss']):.6f} vel {float(l0['vel_loss']):.6f}")
    for mode, g in (("cur", gelu_tanh), ("cur", F.gelu), ("fusedln", gelu_tanh), ("fp32res", gelu_tanh), ("fp32res", F.gelu)):
        t0 = time.time()
        got, hs = full(sd, audio, oh, tp, fps, mode, g)
        l = orm.faceformer_loss(got, gt)
        err = float((got - want).abs().max())
        rel = {k: abs(float(l[k]) - float(l0[k])) / abs(float(l0[k])) for k in l}
        print(f"{mode:8s} gelu={'tanh' if g is gelu_tanh else 'erf '}: max vertex err {err:.3e} cm ({100*err/off:.2f} % of max offset), enc err {float((hs-hs0).abs().max()):.3e} rms {float((hs-hs0).pow(2).mean().sqrt()):.3e}; "
              f"loss rel {rel['loss']:.2e} rec {rel['rec_loss']:.2e} vel {rel['vel_loss']:.2e}  ({time.time()-t0:.1f}s)")
