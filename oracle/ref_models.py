"""TEST INFRASTRUCTURE -- CPU fp32 restatement (oracle) of the reference's audio->mesh hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
the product path (audio2face-pytorch_b200/) never does and fails loudly without its CUDA library.

Every function restates, in plain torch fp32 ops on the CPU, what the cited reference lines compute, operating on a
state_dict with the reference's key names.  The arithmetic of FaceFormer's audio encoder lives in a third-party
dependency that is NOT under /root/reference and is NOT pinned by ref:requirements.txt: HuggingFace `transformers`
Wav2Vec2Model (installed here: transformers 5.5.0; cited as HF:<file>:<line> =
site-packages/transformers/models/wav2vec2/<file>), and torch.nn.TransformerDecoderLayer (torch 2.11.0).

PINNING STATUS: the reference has no tests, golden vectors or checkpoints ("parity unpinned" by the reference
itself).  This oracle is instead pinned against OUTPUTS OF THE LIVE REFERENCE MODULES run in the build
container: tests/golden/make_golden.py imports /root/reference (and HF/torch), loads oracle.weights state_dicts into
the reference classes with strict=True, and stores sub-sampled outputs under tests/golden/*.npz;
tests/test_oracle_golden.py checks this restatement against those fixtures.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------------------
# VOCA  (ref:src/model/voca.py:38-49)
# ------------------------------------------------------------------------------------------------------------
def voca_forward(sd: SD, x: torch.Tensor, one_hot: torch.Tensor, template: torch.Tensor) -> torch.Tensor:
    bs = x.size(0)
    one_hot = one_hot[:, :8]                                            # voca.py:40
    emb = one_hot.repeat(1, 16).view(bs, 1, -1, 16)                     # voca.py:41 (a tiling, not a broadcast)
    h = torch.cat((x.unsqueeze(1), emb), 2)                             # voca.py:42-44 -> [bs,1,37,16]
    h = h.permute(0, 2, 3, 1)                                           # voca.py:45 -> [bs,37,16,1]
    for idx in (0, 2, 4, 6):                                            # voca.py:19-28
        h = F.relu(F.conv2d(h, sd[f"time_conv.{idx}.weight"], sd[f"time_conv.{idx}.bias"], stride=(2, 1), padding=(1, 0)))
    h = torch.cat([h.reshape(bs, -1), one_hot], 1)                      # voca.py:47
    h = F.linear(h, sd["decoder.0.weight"], sd["decoder.0.bias"])       # voca.py:30-36
    h = torch.tanh(F.linear(h, sd["decoder.1.weight"], sd["decoder.1.bias"]))
    h = F.linear(h, sd["decoder.3.weight"], sd["decoder.3.bias"])
    h = F.linear(h, sd["decoder.4.weight"], sd["decoder.4.bias"])
    return h.view(bs, -1, 3) + template                                 # voca.py:49


# ------------------------------------------------------------------------------------------------------------
# Audio2Mesh  (ref:src/model/audio2face.py:57-66)
# ------------------------------------------------------------------------------------------------------------
def _bn(sd: SD, prefix: str, h: torch.Tensor, train: bool, running: Optional[SD] = None) -> torch.Tensor:
    # nn.BatchNorm2d defaults eps=1e-5, momentum=0.1; eval: running stats; train: biased batch stats.  When `running`
    # is given (train mode) its <prefix>.running_mean / running_var tensors are updated IN PLACE exactly as
    # nn.BatchNorm2d does (momentum 0.1, unbiased variance) and <prefix>.num_batches_tracked is incremented.
    if train and running is not None:
        out = F.batch_norm(h, running[prefix + ".running_mean"], running[prefix + ".running_var"], sd[prefix + ".weight"],
                           sd[prefix + ".bias"], training=True, momentum=0.1, eps=1e-5)
        running[prefix + ".num_batches_tracked"] += 1
        return out
    return F.batch_norm(h, None if train else sd[prefix + ".running_mean"], None if train else sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], training=train, momentum=0.0, eps=1e-5)


def audio2mesh_forward(sd: SD, x: torch.Tensor, one_hot: torch.Tensor, template: torch.Tensor,
                       train_bn: bool = False, running: Optional[SD] = None) -> torch.Tensor:
    bs = x.size(0)
    emb = one_hot.repeat(1, 32).view(bs, 1, -1, 32)                     # audio2face.py:59
    h = torch.cat((x.unsqueeze(1), emb), 2)                             # audio2face.py:60-62 -> [bs,1,64,32]
    for i in range(5):                                                  # audio2face.py:13-29 conv -> BN -> ReLU
        h = F.conv2d(h, sd[f"analysis_net.{3 * i}.weight"], sd[f"analysis_net.{3 * i}.bias"], stride=(1, 2), padding=(0, 1))
        h = F.relu(_bn(sd, f"analysis_net.{3 * i + 1}", h, train_bn, running))
    for conv_idx, bn_idx in ((0, 1), (3, 4), (6, 7)):                   # audio2face.py:31-40 conv -> BN -> ReLU
        h = F.conv2d(h, sd[f"articulation_net.{conv_idx}.weight"], sd[f"articulation_net.{conv_idx}.bias"],
                     stride=(2, 1), padding=(1, 0))
        h = F.relu(_bn(sd, f"articulation_net.{bn_idx}", h, train_bn, running))
    h = _bn(sd, "articulation_net.9", h, train_bn, running)                      # audio2face.py:41-43 BN -> conv -> ReLU
    h = F.relu(F.conv2d(h, sd["articulation_net.10.weight"], sd["articulation_net.10.bias"], stride=(2, 1), padding=(1, 0)))
    h = _bn(sd, "articulation_net.12", h, train_bn, running)                     # audio2face.py:44-46
    h = F.relu(F.conv2d(h, sd["articulation_net.13.weight"], sd["articulation_net.13.bias"], stride=(4, 1)))
    h = h.view(bs, -1)                                                  # audio2face.py:64
    h = torch.cat((h, one_hot), 1)                                      # audio2face.py:65
    h = F.linear(h, sd["output_net.0.weight"], sd["output_net.0.bias"])   # audio2face.py:49-55
    h = torch.tanh(F.linear(h, sd["output_net.1.weight"], sd["output_net.1.bias"]))
    h = F.linear(h, sd["output_net.3.weight"], sd["output_net.3.bias"])
    h = F.linear(h, sd["output_net.4.weight"], sd["output_net.4.bias"])
    return h.view(bs, -1, 3) + template                                 # audio2face.py:66


def lstm_forward(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh) -> torch.Tensor:
    """torch.nn.LSTM(batch_first=True, num_layers=1, unidirectional) forward with zero initial state, written out
    (gate order i, f, g, o; torch/nn/modules/rnn.py): x [B,T,I] -> [B,T,H]."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = []
    xp = F.linear(x, w_ih, b_ih + b_hh)
    for t in range(T):
        g = xp[:, t] + F.linear(h, w_hh)
        i, f, gg, o = g.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out.append(h)
    return torch.stack(out, 1)


def song2face_forward(sd: SD, x: torch.Tensor, one_hot: torch.Tensor, template: torch.Tensor) -> torch.Tensor:
    """ref:src/model/song2face.py:63-76, eval mode (running-stat BatchNorm)."""
    bs = x.size(0)
    emb = one_hot.repeat(1, 32).view(bs, 1, -1, 32)                     # song2face.py:65
    h = torch.cat((x.unsqueeze(1), emb), 2)                             # :66-67 -> [bs,1,64,32]
    for i, (kw, pad) in enumerate(((5, 2), (5, 2), (3, 1), (3, 1), (3, 1))):        # :33-39 conv -> BN -> ReLU
        p = f"vocal_encoder_nn.{i}."
        h = F.conv2d(h, sd[p + "0.weight"], sd[p + "0.bias"], stride=(1, 2), padding=(0, pad))
        h = F.relu(_bn(sd, p + "1", h, False))
    h = h.squeeze(3)                                                    # :68 -> [bs, 256, 64]
    for name in ("vocal_encoder_lstm1", "vocal_encoder_lstm2"):         # :69-70: 256 steps of 64 / 256 features
        h = lstm_forward(h, sd[name + ".weight_ih_l0"], sd[name + ".weight_hh_l0"], sd[name + ".bias_ih_l0"],
                         sd[name + ".bias_hh_l0"])
    h = F.interpolate(h.unsqueeze(3), size=(32, 1), mode="bilinear")    # :71-72 -> [bs, 256, 32, 1]
    for i in range(4):                                                  # :54-59
        p = f"regression_net.{i}."
        h = F.conv2d(h, sd[p + "0.weight"], sd[p + "0.bias"], stride=(2, 1), padding=(1 if i < 3 else 0, 0))
        if i < 3:
            h = _bn(sd, p + "1", h, False)
        h = F.relu(h)
    h = h.squeeze(3).squeeze(2)                                         # :74
    h = torch.cat((h, one_hot), 1)                                      # :75
    h = F.linear(h, sd["output_net.0.weight"], sd["output_net.0.bias"])
    h = torch.tanh(F.linear(h, sd["output_net.1.weight"], sd["output_net.1.bias"]))
    h = F.linear(h, sd["output_net.3.weight"], sd["output_net.3.bias"])
    h = F.linear(h, sd["output_net.4.weight"], sd["output_net.4.bias"])
    return h.view(bs, -1, 3) + template                                 # :76


# ------------------------------------------------------------------------------------------------------------
# Losses  (ref:src/loss/loss.py)
# ------------------------------------------------------------------------------------------------------------
def voca_loss(pred: torch.Tensor, gt: torch.Tensor, k_rec: float = 1.0, k_vel: float = 10.0) -> Dict[str, torch.Tensor]:
    bs = pred.shape[0]                                                  # loss.py:42-46
    gt = gt.reshape(bs, -1, 3)
    pred = pred.reshape(bs, -1, 3)
    nv = pred.shape[1]
    rec = torch.mean(torch.sum((pred - gt) ** 2, dim=2))                # loss.py:29-30
    p = pred.reshape(-1, 2, nv, 3)                                      # loss.py:32-40
    g = gt.reshape(-1, 2, nv, 3)
    vel = torch.mean(torch.sum(((p[:, 1] - p[:, 0]) - (g[:, 1] - g[:, 0])) ** 2, dim=2))
    return {"loss": rec * k_rec + vel * k_vel, "rec_loss": rec, "vel_loss": vel}   # loss.py:51-55


def faceformer_loss(pred: torch.Tensor, gt: torch.Tensor) -> Dict[str, torch.Tensor]:
    gt = gt.squeeze(0)                                                  # loss.py:9-17
    pred = pred.squeeze(0)
    if gt.shape[0] % 2 != 0:
        gt = gt[:-1]
        pred = pred[:-1]
    return voca_loss(pred, gt)


def mse_error(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """ref:src/model/lightning_model.py:119-125"""
    pred = pred.reshape(-1, 5023 * 3)
    gt = gt.reshape(-1, 5023 * 3)
    return torch.mean(torch.mean((pred - gt) ** 2, dim=1))


# ------------------------------------------------------------------------------------------------------------
# FaceFormer pieces
# ------------------------------------------------------------------------------------------------------------
def processor_normalize(audio_1d: torch.Tensor) -> torch.Tensor:
    """Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm (HF:feature_extraction_wav2vec2.py:78-98), numpy fp32 on the
    host in the reference (ref:src/model/faceformer.py:142-144): (x - mean) / sqrt(var + 1e-7), population variance."""
    x = audio_1d.detach().cpu().numpy()
    y = (x - x.mean()) / (x.var() + 1e-7) ** 0.5
    return torch.from_numpy(y.astype("float32"))


def init_biased_mask(n_head: int, max_seq_len: int, period: int) -> torch.Tensor:
    """Closed form of ref:src/model/faceformer.py:22-54 (checked equal to the reference function in
    tests/golden/make_golden.py): mask[h,i,j] = -slope_h*floor((i-j)/period) for j<=i, -inf above the diagonal,
    slopes 2^-2, 2^-4, 2^-6, 2^-8 for 4 heads."""
    assert n_head == 4
    slopes = torch.tensor([2.0 ** (-2 * (h + 1)) for h in range(n_head)])
    i = torch.arange(max_seq_len).unsqueeze(1)
    j = torch.arange(max_seq_len).unsqueeze(0)
    steps = torch.div(i - j, period, rounding_mode="floor").float()
    alibi = -slopes.view(-1, 1, 1) * steps.unsqueeze(0)
    causal = torch.zeros(max_seq_len, max_seq_len).masked_fill(j > i, float("-inf"))
    return torch.where((j > i).unsqueeze(0), causal.unsqueeze(0).expand(n_head, -1, -1), alibi + 0.0)


def enc_dec_mask(T: int, S: int) -> torch.Tensor:
    """ref:src/model/faceformer.py:58-66, dataset == "vocaset": only the diagonal is visible (True = masked)."""
    mask = torch.ones(T, S, dtype=torch.bool)
    idx = torch.arange(min(T, S))
    mask[idx, idx] = False
    return mask


def feature_extractor(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """Wav2Vec2FeatureEncoder (HF:modeling_wav2vec2.py:409-419; layers :254-272 and :302-323). x: [B,N] -> [B,512,L]."""
    p = "audio_encoder.feature_extractor.conv_layers."
    h = x[:, None]
    h = F.conv1d(h, sd[p + "0.conv.weight"], None, stride=5)
    h = F.group_norm(h, 512, sd[p + "0.layer_norm.weight"], sd[p + "0.layer_norm.bias"], eps=1e-5)
    h = F.gelu(h)
    for i in range(1, 7):
        h = F.gelu(F.conv1d(h, sd[p + f"{i}.conv.weight"], None, stride=2))
    return h


def linear_interpolation(features: torch.Tensor, output_len: int) -> torch.Tensor:
    """ref:src/model/wav2vec.py:76-84 with output_len given (input_fps/output_fps are then unused)."""
    features = features.transpose(1, 2)
    out = F.interpolate(features, size=output_len, align_corners=True, mode="linear")
    return out.transpose(1, 2)


def feature_projection(sd: SD, h: torch.Tensor) -> torch.Tensor:
    """HF:modeling_wav2vec2.py:429-434 (dropout inactive in eval)."""
    p = "audio_encoder.feature_projection."
    h = F.layer_norm(h, (512,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)
    return F.linear(h, sd[p + "projection.weight"], sd[p + "projection.bias"])


def pos_conv_weight(sd: SD) -> torch.Tensor:
    """weight_norm(dim=2): w = v * (g / ||v||_{dims 0,1})  (torch._weight_norm)."""
    p = "audio_encoder.encoder.pos_conv_embed.conv.parametrizations.weight."
    g, v = sd[p + "original0"], sd[p + "original1"]
    return torch._weight_norm(v, g, 2)


def encoder(sd: SD, h: torch.Tensor) -> torch.Tensor:
    """Wav2Vec2Encoder.forward (HF:modeling_wav2vec2.py:668-727), post-LN layers (:576-609), eager attention
    (:438-463), eval mode, attention_mask=None."""
    e = "audio_encoder.encoder."
    pos = F.conv1d(h.transpose(1, 2), pos_conv_weight(sd), sd[e + "pos_conv_embed.conv.bias"], padding=64, groups=16)
    pos = F.gelu(pos[:, :, :-1]).transpose(1, 2)                        # SamePad drops the last step (:371-379)
    h = h + pos
    h = F.layer_norm(h, (768,), sd[e + "layer_norm.weight"], sd[e + "layer_norm.bias"], 1e-5)
    B, T, _ = h.shape
    for l in range(12):
        p = e + f"layers.{l}."
        q = F.linear(h, sd[p + "attention.q_proj.weight"], sd[p + "attention.q_proj.bias"]).view(B, T, 12, 64).transpose(1, 2)
        k = F.linear(h, sd[p + "attention.k_proj.weight"], sd[p + "attention.k_proj.bias"]).view(B, T, 12, 64).transpose(1, 2)
        v = F.linear(h, sd[p + "attention.v_proj.weight"], sd[p + "attention.v_proj.bias"]).view(B, T, 12, 64).transpose(1, 2)
        w = torch.softmax(torch.matmul(q, k.transpose(2, 3)) * (64 ** -0.5), dim=-1)
        a = torch.matmul(w, v).transpose(1, 2).reshape(B, T, 768)
        a = F.linear(a, sd[p + "attention.out_proj.weight"], sd[p + "attention.out_proj.bias"])
        h = F.layer_norm(h + a, (768,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)
        f = F.gelu(F.linear(h, sd[p + "feed_forward.intermediate_dense.weight"], sd[p + "feed_forward.intermediate_dense.bias"]))
        f = F.linear(f, sd[p + "feed_forward.output_dense.weight"], sd[p + "feed_forward.output_dense.bias"])
        h = F.layer_norm(h + f, (768,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)
    return h


def spec_augment_time_mask(batch: int, frames: int, mask_prob: float = 0.05, mask_length: int = 10,
                           min_masks: int = 2, rng=np.random) -> np.ndarray:
    """`_compute_mask_indices((B,T), mask_time_prob, mask_time_length, None, min_masks=2)` of
    ref:src/model/wav2vec.py:25-72 as called at :153-159 (HF Wav2Vec2Config defaults 0.05 / 10), for
    attention_mask=None.  Draws from `rng` (numpy's global generator by default) in the reference's order:
    rand() once (:37), choice(sz-min_len, n) per row (:59), choice(idc, min_len) per over-long row (:70-71)."""
    num = max(min_masks, int(mask_prob * frames / float(mask_length) + rng.rand()))      # :37-38
    rows = []
    for _ in range(batch):
        min_len = mask_length                                                            # :51-55 (all spans equal)
        if frames - min_len <= num:
            min_len = frames - num - 1                                                   # :56-57
        start = rng.choice(frames - min_len, num, replace=False)                         # :59
        idc = np.asarray([start[j] + o for j in range(num) for o in range(mask_length)]) # :60-66
        rows.append(np.unique(idc[idc < frames]))                                        # :67
    shortest = min(len(r) for r in rows)                                                 # :69
    mask = np.full((batch, frames), False)
    for i, r in enumerate(rows):
        if len(r) > shortest:
            r = rng.choice(r, shortest, replace=False)                                   # :71-72
        mask[i, r] = True
    return mask


def audio_encoder(sd: SD, audio_norm: torch.Tensor, frame_num: int, spec_mask=None) -> torch.Tensor:
    """ref:src/model/wav2vec.py:91-187 for dataset == "vocaset"; eval mode when spec_mask is None, otherwise the
    training branch :149-162 with the given bool [B,T] time mask (dropout / LayerDrop stay off)."""
    h = feature_extractor(sd, audio_norm).transpose(1, 2)               # wav2vec.py:116-117
    h = linear_interpolation(h, frame_num)                              # wav2vec.py:125-128
    h = feature_projection(sd, h)                                       # wav2vec.py:147
    if spec_mask is not None:
        m = torch.as_tensor(np.asarray(spec_mask), dtype=torch.bool)[..., None]
        h = torch.where(m, sd["audio_encoder.masked_spec_embed"].to(h.dtype), h)        # wav2vec.py:159-161
    return encoder(sd, h)                                               # wav2vec.py:174-180


def _mha(x_q, x_kv, w_in, b_in, w_out, b_out, n_head, attn_mask_float=None, attn_mask_bool=None):
    """torch F.multi_head_attention_forward math (packed in_proj: rows q | k | v), batch_first inputs [1,t,64]."""
    E = x_q.shape[-1]
    hd = E // n_head
    q = F.linear(x_q, w_in[:E], b_in[:E])
    k = F.linear(x_kv, w_in[E:2 * E], b_in[E:2 * E])
    v = F.linear(x_kv, w_in[2 * E:], b_in[2 * E:])
    t, s = q.shape[1], k.shape[1]
    q = q.view(t, n_head, hd).transpose(0, 1)
    k = k.view(s, n_head, hd).transpose(0, 1)
    v = v.view(s, n_head, hd).transpose(0, 1)
    scores = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(hd)
    if attn_mask_float is not None:
        scores = scores + attn_mask_float
    if attn_mask_bool is not None:
        scores = scores.masked_fill(attn_mask_bool, float("-inf"))
    a = torch.matmul(torch.softmax(scores, dim=-1), v)                  # [h,t,hd]
    a = a.transpose(0, 1).reshape(1, t, E)
    return F.linear(a, w_out, b_out)


def decoder_layer(sd: SD, x: torch.Tensor, memory: torch.Tensor, tgt_mask: torch.Tensor, memory_mask: torch.Tensor):
    """nn.TransformerDecoderLayer(d_model=64, nhead=4, dim_feedforward=128, batch_first=True), norm_first=False,
    ReLU, eval mode (torch/nn/modules/transformer.py:1145-1156, :1158-1199) as built at ref:faceformer.py:121-127."""
    t = "transformer_decoder.layers.0."
    sa = _mha(x, x, sd[t + "self_attn.in_proj_weight"], sd[t + "self_attn.in_proj_bias"],
              sd[t + "self_attn.out_proj.weight"], sd[t + "self_attn.out_proj.bias"], 4, attn_mask_float=tgt_mask)
    x = F.layer_norm(x + sa, (64,), sd[t + "norm1.weight"], sd[t + "norm1.bias"], 1e-5)
    ca = _mha(x, memory, sd[t + "multihead_attn.in_proj_weight"], sd[t + "multihead_attn.in_proj_bias"],
              sd[t + "multihead_attn.out_proj.weight"], sd[t + "multihead_attn.out_proj.bias"], 4,
              attn_mask_bool=memory_mask)
    x = F.layer_norm(x + ca, (64,), sd[t + "norm2.weight"], sd[t + "norm2.bias"], 1e-5)
    ff = F.linear(F.relu(F.linear(x, sd[t + "linear1.weight"], sd[t + "linear1.bias"])),
                  sd[t + "linear2.weight"], sd[t + "linear2.bias"])
    return F.layer_norm(x + ff, (64,), sd[t + "norm3.weight"], sd[t + "norm3.bias"], 1e-5)


def faceformer_decode(sd: SD, hidden_states: torch.Tensor, one_hot: torch.Tensor, frame_num: int,
                      period: int = 60) -> torch.Tensor:
    """The autoregressive loop of ref:src/model/faceformer.py:148,154-185 restated literally (the whole prefix and
    the 64->15069 head are recomputed every step).  Returns vertice_out [1,T,V3] before the template add.
    For frame_num > 600 the biased mask / PPE are rebuilt at the needed length (closed forms)."""
    max_len = max(600, frame_num)
    biased_mask = init_biased_mask(4, max_len, period)
    pe = sd["PPE.pe"]
    if pe.shape[1] < frame_num:
        reps = frame_num // period + 1
        pe = pe[:, :period].repeat(1, reps, 1)
    obj_embedding = F.linear(one_hot, sd["obj_vector.weight"])          # faceformer.py:148
    style_emb = obj_embedding.unsqueeze(1)
    vertice_emb = style_emb
    vertice_out = None
    for i in range(frame_num):
        vertice_input = vertice_emb + pe[:, : vertice_emb.shape[1], :]  # faceformer.py:155-160 (PPE, dropout off)
        t = vertice_input.shape[1]
        tgt_mask = biased_mask[:, :t, :t]                               # faceformer.py:162-167
        memory_mask = enc_dec_mask(t, hidden_states.shape[1])           # faceformer.py:168-173
        dec = decoder_layer(sd, vertice_input, hidden_states, tgt_mask, memory_mask)   # faceformer.py:174-179
        vertice_out = F.linear(dec, sd["vertice_map_r.weight"], sd["vertice_map_r.bias"])   # faceformer.py:181
        new_output = F.linear(vertice_out[:, -1, :], sd["vertice_map.weight"], sd["vertice_map.bias"]).unsqueeze(1)
        new_output = new_output + style_emb                             # faceformer.py:183-184
        vertice_emb = torch.cat((vertice_emb, new_output), 1)           # faceformer.py:185
    return vertice_out


def faceformer_forward(sd: SD, audio: torch.Tensor, one_hot: torch.Tensor, template: torch.Tensor,
                       fps: int = 60, return_parts: bool = False, spec_mask=None):
    """ref:src/model/faceformer.py:139-188 for one utterance: audio [1,N] raw 16 kHz, one_hot [1,n], template
    [1,5023,3] -> [1,T,5023,3].  fps=60 is the reference's hard-coded rate (faceformer.py:141); other values are the
    documented extension: only frame_num changes."""
    frame_num = audio.shape[1] * fps // 16000                           # faceformer.py:141
    audio_n = processor_normalize(audio.squeeze(0))[None]               # faceformer.py:142-144
    template = template.reshape(1, 1, -1)                               # faceformer.py:147
    hs = audio_encoder(sd, audio_n, frame_num, spec_mask)               # faceformer.py:149-151
    memory = F.linear(hs, sd["audio_feature_map.weight"], sd["audio_feature_map.bias"])   # faceformer.py:152
    vertice_out = faceformer_decode(sd, memory, one_hot, frame_num)
    out = (vertice_out + template).view(1, frame_num, -1, 3)            # faceformer.py:187-188
    if return_parts:
        return out, {"encoder": hs, "memory": memory}
    return out


def faceformer_forward_batch(sd: SD, audio: torch.Tensor, one_hot: torch.Tensor, template: torch.Tensor,
                             fps: int = 60) -> torch.Tensor:
    """Batch extension used by the B200 path: per-utterance reference semantics, stacked."""
    outs = [faceformer_forward(sd, audio[b:b + 1], one_hot[b:b + 1], template[b:b + 1], fps) for b in range(audio.shape[0])]
    return torch.cat(outs, 0)
