"""TEST INFRASTRUCTURE (oracle side) -- deterministic random state_dicts for the three models of the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

The reference ships no checkpoint and the pretrained wav2vec2 weights cannot be downloaded (no network), so every
parity check runs on random-init weights (BASELINE.json: "random-init").  The weights are generated here with
numpy's PCG64 keyed by (seed, crc32(parameter name)) -- reproducible bit-for-bit on any machine -- with the exact
state_dict key names / shapes of the reference modules (tests/golden/make_golden.py asserts the
key set and shapes against the live reference modules and loads these tensors into them with strict=True).

Scales are chosen so that activations stay O(1) through the stacks (fan-in scaled), norm layers get non-trivial
affine parameters and BatchNorm gets non-trivial running statistics, and FaceFormer's zero-initialised vertex heads
(ref:src/model/faceformer.py:132-135) are overwritten with N(0, 0.02) -- otherwise the output is exactly the
template and any parity test is vacuous.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import numpy as np
import torch

V3 = 15069


def voca_shapes(n_verts: int = V3) -> "OrderedDict[str, tuple]":
    """ref:src/model/voca.py:19-36"""
    d = OrderedDict()
    for idx, (co, ci) in zip((0, 2, 4, 6), ((32, 37), (32, 32), (64, 32), (64, 64))):
        d[f"time_conv.{idx}.weight"] = (co, ci, 3, 1)
        d[f"time_conv.{idx}.bias"] = (co,)
    for idx, (n, k) in zip((0, 1, 3, 4), ((72, 72), (128, 72), (50, 128), (n_verts, 50))):
        d[f"decoder.{idx}.weight"] = (n, k)
        d[f"decoder.{idx}.bias"] = (n,)
    return d


def audio2mesh_shapes(n_verts: int = V3, n_onehot: int = 12) -> "OrderedDict[str, tuple]":
    """ref:src/model/audio2face.py:13-55"""
    d = OrderedDict()

    def bn(prefix, c):
        d[f"{prefix}.weight"] = (c,)
        d[f"{prefix}.bias"] = (c,)
        d[f"{prefix}.running_mean"] = (c,)
        d[f"{prefix}.running_var"] = (c,)
        d[f"{prefix}.num_batches_tracked"] = ()

    chans = [1, 72, 108, 162, 243, 256]
    for i in range(5):
        d[f"analysis_net.{3 * i}.weight"] = (chans[i + 1], chans[i], 1, 3)
        d[f"analysis_net.{3 * i}.bias"] = (chans[i + 1],)
        bn(f"analysis_net.{3 * i + 1}", chans[i + 1])
    # articulation: conv(0) bn(1) relu | conv(3) bn(4) relu | conv(6) bn(7) relu | bn(9) conv(10) relu | bn(12) conv(13) relu
    for conv_idx, bn_idx, kh in ((0, 1, 3), (3, 4, 3), (6, 7, 3)):
        d[f"articulation_net.{conv_idx}.weight"] = (256, 256, kh, 1)
        d[f"articulation_net.{conv_idx}.bias"] = (256,)
        bn(f"articulation_net.{bn_idx}", 256)
    for bn_idx, conv_idx, kh in ((9, 10, 3), (12, 13, 4)):
        bn(f"articulation_net.{bn_idx}", 256)
        d[f"articulation_net.{conv_idx}.weight"] = (256, 256, kh, 1)
        d[f"articulation_net.{conv_idx}.bias"] = (256,)
    for idx, (n, k) in zip((0, 1, 3, 4), ((72, 256 + n_onehot), (128, 72), (50, 128), (n_verts, 50))):
        d[f"output_net.{idx}.weight"] = (n, k)
        d[f"output_net.{idx}.bias"] = (n,)
    return d


def song2face_shapes(n_verts: int = V3, n_onehot: int = 12) -> "OrderedDict[str, tuple]":
    """ref:src/model/song2face.py:32-61"""
    d = OrderedDict()

    def conv_bn(prefix, ci, co, kh, kw, bn=True):
        d[f"{prefix}.0.weight"] = (co, ci, kh, kw)
        d[f"{prefix}.0.bias"] = (co,)
        if bn:
            for leaf, shp in (("weight", (co,)), ("bias", (co,)), ("running_mean", (co,)), ("running_var", (co,)),
                              ("num_batches_tracked", ())):
                d[f"{prefix}.1.{leaf}"] = shp

    chans = [1, 72, 108, 162, 243, 256]
    for i, kw in enumerate((5, 5, 3, 3, 3)):
        conv_bn(f"vocal_encoder_nn.{i}", chans[i], chans[i + 1], 1, kw)
    for name, insz in (("vocal_encoder_lstm1", 64), ("vocal_encoder_lstm2", 256)):
        d[f"{name}.weight_ih_l0"] = (1024, insz)
        d[f"{name}.weight_hh_l0"] = (1024, 256)
        d[f"{name}.bias_ih_l0"] = (1024,)
        d[f"{name}.bias_hh_l0"] = (1024,)
    for idx, (n, k) in zip((0, 1, 3, 4), ((72, 256 + n_onehot), (128, 72), (50, 128), (n_verts, 50))):
        d[f"output_net.{idx}.weight"] = (n, k)
        d[f"output_net.{idx}.bias"] = (n,)
    for i in range(4):
        conv_bn(f"regression_net.{i}", 256, 256, 3, 1, bn=i < 3)
    return d


def faceformer_shapes(n_verts: int = V3, n_onehot: int = 12) -> "OrderedDict[str, tuple]":
    """ref:src/model/faceformer.py:91-135 + transformers Wav2Vec2Config() defaults (base architecture)."""
    d = OrderedDict()
    ae = "audio_encoder."
    d[ae + "masked_spec_embed"] = (768,)
    fe = ae + "feature_extractor.conv_layers."
    d[fe + "0.conv.weight"] = (512, 1, 10)
    d[fe + "0.layer_norm.weight"] = (512,)
    d[fe + "0.layer_norm.bias"] = (512,)
    for i, k in zip(range(1, 7), (3, 3, 3, 3, 2, 2)):
        d[fe + f"{i}.conv.weight"] = (512, 512, k)
    fp = ae + "feature_projection."
    d[fp + "layer_norm.weight"] = (512,)
    d[fp + "layer_norm.bias"] = (512,)
    d[fp + "projection.weight"] = (768, 512)
    d[fp + "projection.bias"] = (768,)
    enc = ae + "encoder."
    d[enc + "pos_conv_embed.conv.bias"] = (768,)
    d[enc + "pos_conv_embed.conv.parametrizations.weight.original0"] = (1, 1, 128)
    d[enc + "pos_conv_embed.conv.parametrizations.weight.original1"] = (768, 48, 128)
    d[enc + "layer_norm.weight"] = (768,)
    d[enc + "layer_norm.bias"] = (768,)
    for l in range(12):
        p = enc + f"layers.{l}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            d[p + f"attention.{nm}.weight"] = (768, 768)
            d[p + f"attention.{nm}.bias"] = (768,)
        d[p + "layer_norm.weight"] = (768,)
        d[p + "layer_norm.bias"] = (768,)
        d[p + "feed_forward.intermediate_dense.weight"] = (3072, 768)
        d[p + "feed_forward.intermediate_dense.bias"] = (3072,)
        d[p + "feed_forward.output_dense.weight"] = (768, 3072)
        d[p + "feed_forward.output_dense.bias"] = (768,)
        d[p + "final_layer_norm.weight"] = (768,)
        d[p + "final_layer_norm.bias"] = (768,)
    d["audio_feature_map.weight"] = (64, 768)
    d["audio_feature_map.bias"] = (64,)
    d["vertice_map.weight"] = (64, n_verts)
    d["vertice_map.bias"] = (64,)
    d["PPE.pe"] = (1, 660, 64)
    t = "transformer_decoder.layers.0."
    for att in ("self_attn", "multihead_attn"):
        d[t + att + ".in_proj_weight"] = (192, 64)
        d[t + att + ".in_proj_bias"] = (192,)
        d[t + att + ".out_proj.weight"] = (64, 64)
        d[t + att + ".out_proj.bias"] = (64,)
    d[t + "linear1.weight"] = (128, 64)
    d[t + "linear1.bias"] = (128,)
    d[t + "linear2.weight"] = (64, 128)
    d[t + "linear2.bias"] = (64,)
    for i in (1, 2, 3):
        d[t + f"norm{i}.weight"] = (64,)
        d[t + f"norm{i}.bias"] = (64,)
    d["vertice_map_r.weight"] = (n_verts, 64)
    d["vertice_map_r.bias"] = (n_verts,)
    d["obj_vector.weight"] = (64, n_onehot)
    return d


def ppe_table(d_model: int = 64, period: int = 60, max_seq_len: int = 600) -> torch.Tensor:
    """PPE.pe exactly as ref:src/model/faceformer.py:74-84 builds it."""
    pe = torch.zeros(period, d_model)
    position = torch.arange(0, period, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    pe = pe.unsqueeze(0)
    return pe.repeat(1, (max_seq_len // period) + 1, 1)


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def _fill(name: str, shape: tuple, seed: int) -> torch.Tensor:
    r = _rng(seed, name)
    leaf = name.rsplit(".", 1)[-1]
    is_norm = any(s in name for s in ("layer_norm", "norm1", "norm2", "norm3", "final_layer_norm")) or (
        ("analysis_net" in name or "articulation_net" in name) and len(shape) == 1 and leaf != "bias"
    )
    if leaf == "num_batches_tracked":
        return torch.tensor(7, dtype=torch.int64)
    if leaf == "running_mean":
        return torch.from_numpy((0.1 * r.standard_normal(shape)).astype(np.float32))
    if leaf == "running_var":
        return torch.from_numpy(r.uniform(0.5, 1.5, shape).astype(np.float32))
    if name == "PPE.pe":
        return ppe_table()
    if name.endswith("parametrizations.weight.original0"):       # weight-norm gain g, per tap
        return torch.from_numpy(r.uniform(0.5, 1.5, shape).astype(np.float32))
    if name.startswith(("vertice_map", "obj_vector")) or name.endswith("masked_spec_embed"):
        std = 0.02 if not name.startswith("obj_vector") else 0.3
        return torch.from_numpy((std * r.standard_normal(shape)).astype(np.float32))
    if len(shape) == 1:
        # BatchNorm / LayerNorm / GroupNorm affine weights, or biases
        bn_affine = ("analysis_net" in name or "articulation_net" in name)
        if leaf == "weight" and (is_norm or bn_affine):
            return torch.from_numpy(r.uniform(0.5, 1.5, shape).astype(np.float32))
        if leaf == "bias" and is_norm:
            return torch.from_numpy((0.1 * r.standard_normal(shape)).astype(np.float32))
        return torch.from_numpy((0.05 * r.standard_normal(shape)).astype(np.float32))
    fan_in = int(np.prod(shape[1:]))
    gain = 1.4 if ("conv_layers" in name or "time_conv" in name or "analysis_net" in name or "vocal_encoder_nn" in name
                   or "regression_net" in name
                   or "articulation_net" in name or "intermediate_dense" in name or "linear1" in name) else 1.0
    std = gain / math.sqrt(max(fan_in, 1))
    return torch.from_numpy((std * r.standard_normal(shape)).astype(np.float32))


def _is_bn_vector(name: str, shapes) -> bool:
    base = name.rsplit(".", 1)[0]
    return (base + ".running_mean") in shapes


def make_state_dict(model: str, seed: int = 0, n_verts: int = V3, n_onehot: int = 12) -> "OrderedDict[str, torch.Tensor]":
    shapes = {"voca": voca_shapes, "audio2mesh": audio2mesh_shapes, "faceformer": faceformer_shapes,
              "song2face": song2face_shapes}[model]
    shp = shapes(n_verts) if model == "voca" else shapes(n_verts, n_onehot)
    sd = OrderedDict()
    for name, shape in shp.items():
        if model in ("audio2mesh", "song2face") and _is_bn_vector(name, shp) and name.endswith((".weight", ".bias")):
            r = _rng(seed, name)
            if name.endswith(".weight"):
                sd[name] = torch.from_numpy(r.uniform(0.5, 1.5, shape).astype(np.float32))
            else:
                sd[name] = torch.from_numpy((0.1 * r.standard_normal(shape)).astype(np.float32))
        else:
            sd[name] = _fill(name, shape, seed)
    return sd
