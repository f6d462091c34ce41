"""CPU oracle for the TubeR inference forward path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32) *restatement* of the reference's forward
algorithm, written functionally over a flat ``{state_dict name: tensor}`` mapping.
It exists to check the CUDA path; it is never imported by the product package
(``tubelet-transformer_b200/``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Where the arithmetic lives: the reference has no kernels of its own -- every op is a
stock ``torch.nn`` module (PyTorch is a third-party dependency, not vendored; the
reference README pins it only as "Torch 1.12 + CUDA 11.3", README.md:41).  The
restatement therefore uses the same primitive torch ops (conv3d, batch_norm, linear,
softmax, layer_norm) and re-derives everything the reference *composes* from them:
multi-head attention, the encoder/decoder wiring, the class branch, the temporal pool,
the 3-D sine position code.

Parity pin: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the reference itself,
run in the build container by ``oracle/make_golden.py`` (which imports
``/root/reference`` unmodified) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks this file against every one of those fixtures.

Citations (``file:line``) are relative to the reference repository root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # models/backbones/ir_CSN_152.py:15
LN_EPS = 1e-5          # torch.nn.LayerNorm default, used throughout the reference
STAGE_BLOCKS = {"CSN-152": (3, 8, 36, 3),   # ir_CSN_152.py:202
                "CSN-50": (3, 4, 6, 3)}     # ir_CSN_50.py:204
POOL_DIM = 2048        # backbone_builder.py:49-52 hard-codes d_model=2048 for the decode pool


# ----------------------------------------------------------------------------------------
# configuration helpers
# ----------------------------------------------------------------------------------------
def _m(cfg):
    return cfg.CONFIG.MODEL


def backbone_name(cfg) -> str:
    # backbone_builder.py:31-36: anything that is not 'CSN-152' builds CSN-50
    return "CSN-152" if _m(cfg).BACKBONE_NAME == "CSN-152" else "CSN-50"


def is_ava(cfg) -> bool:
    return cfg.CONFIG.DATA.DATASET_NAME == "ava"      # tuber_ava.py:46,64,70


def num_query_rows(cfg) -> int:
    # tuber_ava.py:46-50
    return _m(cfg).QUERY_NUM if is_ava(cfg) else _m(cfg).QUERY_NUM * _m(cfg).TEMP_LEN


def num_class_out(cfg) -> int:
    # tuber_ava.py:70-73
    return cfg.CONFIG.DATA.NUM_CLASSES if is_ava(cfg) else cfg.CONFIG.DATA.NUM_CLASSES + 1


# ----------------------------------------------------------------------------------------
# state_dict specification (names / shapes of the reference model) and seeded weights
# ----------------------------------------------------------------------------------------
def _bn_spec(prefix: str, c: int) -> List[Tuple[str, tuple, str]]:
    return [(prefix + ".weight", (c,), "bn_weight"), (prefix + ".bias", (c,), "bn_bias"),
            (prefix + ".running_mean", (c,), "bn_mean"), (prefix + ".running_var", (c,), "bn_var"),
            (prefix + ".num_batches_tracked", (), "bn_count")]


def _mha_spec(prefix: str, d: int) -> List[Tuple[str, tuple, str]]:
    return [(prefix + ".in_proj_weight", (3 * d, d), "linear"), (prefix + ".in_proj_bias", (3 * d,), "bias"),
            (prefix + ".out_proj.weight", (d, d), "linear"), (prefix + ".out_proj.bias", (d,), "bias")]


def _linear_spec(prefix: str, i: int, o: int) -> List[Tuple[str, tuple, str]]:
    return [(prefix + ".weight", (o, i), "linear"), (prefix + ".bias", (o,), "bias")]


def _ln_spec(prefix: str, d: int) -> List[Tuple[str, tuple, str]]:
    return [(prefix + ".weight", (d,), "ln_weight"), (prefix + ".bias", (d,), "ln_bias")]


def param_spec(cfg) -> List[Tuple[str, tuple, str]]:
    """(name, shape, kind) for every entry of the reference model's ``state_dict()``.

    Order and names follow module registration in tuber_ava.py:24-81,
    backbone_builder.py:27-57, ir_CSN_152.py:36-68,95-170, transformer.py:17-33,131-149,
    193-211, transformer_layers.py:46-64,170-199,407-422 and criterion.py:485-492;
    ``oracle/make_golden.py`` asserts the set equals the live reference's.
    """
    m = _m(cfg)
    d, ff = m.D_MODEL, m.DIM_FEEDFORWARD
    spec: List[Tuple[str, tuple, str]] = []
    # transformer.{encoder,decoder}
    for i in range(m.ENC_LAYERS):
        p = f"transformer.encoder.layers.{i}"
        spec += _mha_spec(p + ".self_attn", d) + _linear_spec(p + ".linear1", d, ff) + _linear_spec(p + ".linear2", ff, d)
        spec += _ln_spec(p + ".norm1", d) + _ln_spec(p + ".norm2", d)
    for i in range(m.DEC_LAYERS):
        p = f"transformer.decoder.layers.{i}"
        spec += _mha_spec(p + ".self_attn", d) + _mha_spec(p + ".multihead_attn", d)
        spec += _linear_spec(p + ".linear1", d, ff) + _linear_spec(p + ".linear2", ff, d)
        spec += _ln_spec(p + ".norm1", d) + _ln_spec(p + ".norm2", d) + _ln_spec(p + ".norm3", d)
    spec += _ln_spec("transformer.decoder.norm", d)
    spec += [("query_embed.weight", (num_query_rows(cfg), d), "embed")]
    cb = m.DIM_FEEDFORWARD  # backbone.num_channels := DIM_FEEDFORWARD (backbone_builder.py:111)
    spec += [("input_proj.weight", (d, cb, 1, 1, 1), "conv"), ("input_proj.bias", (d,), "bias")]
    spec += [("class_proj.weight", (d, cb, 1, 1, 1), "conv"), ("class_proj.bias", (d,), "bias")]
    # class-branch encoder (tuber_ava.py:60-61: d_model=hidden_dim, nhead 8, ffn 2048 hard-coded)
    p = "encoder.layers.0"
    spec += _mha_spec(p + ".self_attn_t", d) + _mha_spec(p + ".self_attn_s", d)
    spec += _linear_spec(p + ".linear1", 2 * d, 2048) + _linear_spec(p + ".linear2", 2048, d)
    spec += _ln_spec(p + ".norm1_t", d) + _ln_spec(p + ".norm1_s", d) + _ln_spec(p + ".norm2", d)
    spec += _mha_spec("cross_attn", 256)                                   # tuber_ava.py:62
    spec += _linear_spec("class_embed_b", d, 3) if is_ava(cfg) else _linear_spec("class_embed_b", 2048, 2)
    spec += _linear_spec("bbox_embed.layers.0", d, d) + _linear_spec("bbox_embed.layers.1", d, d)
    spec += _linear_spec("bbox_embed.layers.2", d, 4)
    spec += _linear_spec("class_fc", d, num_class_out(cfg))
    # backbone.body
    b = "backbone.body"
    spec += [(b + ".conv1.weight", (64, 3, 3, 7, 7), "conv")] + _bn_spec(b + ".bn1", 64)
    in_planes = 64
    for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), STAGE_BLOCKS[backbone_name(cfg)])):
        out_planes = planes * 4
        for bi in range(nblk):
            p = f"{b}.layer{li + 1}.{bi}"
            cin = in_planes if bi == 0 else out_planes
            spec += [(p + ".conv1.weight", (planes, cin, 1, 1, 1), "conv")] + _bn_spec(p + ".bn1", planes)
            spec += [(p + ".conv3.weight", (planes, 1, 3, 3, 3), "conv")] + _bn_spec(p + ".bn3", planes)
            spec += [(p + ".conv4.weight", (out_planes, planes, 1, 1, 1), "conv")] + _bn_spec(p + ".bn4", out_planes)
            if bi == 0:
                spec += [(p + ".down_sample.0.weight", (out_planes, cin, 1, 1, 1), "conv")]
                spec += _bn_spec(p + ".down_sample.1", out_planes)
        in_planes = out_planes
    if backbone_name(cfg) == "CSN-50":
        # ir_CSN_50.py:137 keeps an (unused) classifier in the state_dict
        spec += _linear_spec(b + ".out_fc", 2048, cfg.CONFIG.DATA.NUM_CLASSES)
    if m.SINGLE_FRAME and m.TEMPORAL_DS_STRATEGY == "decode":
        spec += [("backbone.query_pool.weight", (1, POOL_DIM), "embed")]
        p = "backbone.pool_decoder.layers.0"
        spec += _mha_spec(p + ".self_attn", POOL_DIM) + _mha_spec(p + ".multihead_attn", POOL_DIM)
        spec += _linear_spec(p + ".linear1", POOL_DIM, 2048) + _linear_spec(p + ".linear2", 2048, POOL_DIM)
        spec += _ln_spec(p + ".norm1", POOL_DIM) + _ln_spec(p + ".norm2", POOL_DIM) + _ln_spec(p + ".norm3", POOL_DIM)
        spec += _ln_spec("backbone.pool_decoder.norm", POOL_DIM)
    return spec


def make_state_dict(cfg, seed: int = 0, bn: str = "identity") -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights under the reference's names and shapes.

    Drawn with numpy's PCG64 (stable across platforms and versions) in ``param_spec``
    order, so the same ``(cfg, seed, bn)`` gives bit-identical weights in the build
    container (where the golden fixtures are made) and on the GPU box.
    ``bn='identity'`` is the reference's default BatchNorm state (gamma 1, beta 0, mean 0,
    var 1); ``bn='random'`` is the SURVEY section 8d second variant that makes BN-folding mistakes
    visible (gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1), var~U(.5,1.5)).
    Scales: conv/linear U(+-sqrt(3/fan_in)) (unit gain, keeps the 50-block residual stream
    O(1)); biases U(+-0.1); embeddings N(0,1); LayerNorm gamma~U(.8,1.2), beta~N(0,.05).
    """
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in param_spec(cfg):
        if kind in ("conv", "linear"):
            fan_in = int(np.prod(shape[1:]))
            bound = math.sqrt(3.0 / fan_in)
            a = rng.uniform(-bound, bound, size=shape)
        elif kind == "bias":
            a = rng.uniform(-0.1, 0.1, size=shape)
        elif kind == "embed":
            a = rng.standard_normal(size=shape)
        elif kind == "ln_weight":
            a = rng.uniform(0.8, 1.2, size=shape)
        elif kind == "ln_bias":
            a = 0.05 * rng.standard_normal(size=shape)
        elif kind == "bn_weight":
            a = rng.uniform(0.5, 1.5, size=shape) if bn == "random" else np.ones(shape)
        elif kind == "bn_bias":
            a = 0.1 * rng.standard_normal(size=shape) if bn == "random" else np.zeros(shape)
        elif kind == "bn_mean":
            a = 0.1 * rng.standard_normal(size=shape) if bn == "random" else np.zeros(shape)
        elif kind == "bn_var":
            a = rng.uniform(0.5, 1.5, size=shape) if bn == "random" else np.ones(shape)
        elif kind == "bn_count":
            sd[name] = torch.zeros((), dtype=torch.int64)
            continue
        else:
            raise ValueError(kind)
        sd[name] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


def ltc_param_spec(cfg) -> List[Tuple[str, tuple, str]]:
    """Parameters of the long-term context layer.  NOT IN THE REFERENCE (announced in its README.md:16-18,86,
    never released): defined in this repository, see ``forward(bank=...)``; kept out of ``param_spec`` so that the reference-pinned
    state_dict and every golden fixture stay what they are."""
    d = _m(cfg).D_MODEL
    return _mha_spec("ltc_attn", d) + _ln_spec("ltc_norm", d)


def make_ltc_state_dict(cfg, seed: int = 100) -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights of the long-term context layer (same distributions as ``make_state_dict``)."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in ltc_param_spec(cfg):
        if kind == "linear":
            a = rng.uniform(-math.sqrt(3.0 / shape[1]), math.sqrt(3.0 / shape[1]), size=shape)
        elif kind == "bias":
            a = rng.uniform(-0.1, 0.1, size=shape)
        elif kind == "ln_weight":
            a = rng.uniform(0.8, 1.2, size=shape)
        else:
            a = 0.05 * rng.standard_normal(size=shape)
        sd[name] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


def make_clips(batch: int, t: int, h: int, w: int, seed: int = 2) -> torch.Tensor:
    """Synthetic ImageNet-normalised clips (B,3,T,H,W), N(0,1), PCG64-seeded."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.standard_normal(size=(batch, 3, t, h, w), dtype=np.float32))


IMAGENET_MEAN = (0.485, 0.456, 0.406)      # datasets/ava_frame.py:159-162
IMAGENET_STD = (0.229, 0.224, 0.225)


def make_frames_u8(batch: int, t: int, h: int, w: int, seed: int = 3) -> torch.Tensor:
    """Synthetic decoded RGB frames, uint8 (B,T,H,W,3), PCG64-seeded."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.integers(0, 256, size=(batch, t, h, w, 3), dtype=np.uint8))


def frames_to_clips(frames_u8: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    """uint8 frames (B,T,H,W,3) -> the (B,3,T,H,W) fp32 clips the reference's input pipeline feeds the model:
    per frame ``to_tensor`` (HWC uint8 -> CHW float32, ``.div(255)``; datasets/video_transforms.py:294-296) and
    ``normalize`` (``.sub_(mean).div_(std)`` with fp32 mean / std; :308-314), then ``stack`` + ``permute(1,0,2,3)``
    (datasets/ava_frame.py:71-72).  Same torch ops, batched."""
    x = frames_u8.permute(0, 1, 4, 2, 3).to(torch.float32).div(255)            # (B,T,3,H,W)
    m = torch.as_tensor(mean, dtype=torch.float32).view(1, 1, 3, 1, 1)
    s = torch.as_tensor(std, dtype=torch.float32).view(1, 1, 3, 1, 1)
    x = x.sub_(m).div_(s)
    return x.permute(0, 2, 1, 3, 4).contiguous()


def pad_clips(clips: List[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Zero-pad a ragged list of (3,T,H,W) clips to the batch maximum and build the
    (B,H,W) bool mask, True on padding -- utils/misc.py:385-399."""
    t = max(c.shape[1] for c in clips)
    h = max(c.shape[2] for c in clips)
    w = max(c.shape[3] for c in clips)
    out = torch.zeros((len(clips), 3, t, h, w), dtype=clips[0].dtype)
    mask = torch.ones((len(clips), h, w), dtype=torch.bool)
    for i, c in enumerate(clips):
        out[i, :, : c.shape[1], : c.shape[2], : c.shape[3]] = c
        mask[i, : c.shape[2], : c.shape[3]] = False
    return out, mask


# ----------------------------------------------------------------------------------------
# primitive blocks
# ----------------------------------------------------------------------------------------
def _bn(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    # eval-mode BatchNorm3d(eps=1e-3): ir_CSN_152.py:46,56,64,119,154
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], training=False, eps=BN_EPS)


def _ln(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], LN_EPS)


def mha(sd, prefix: str, q_in: torch.Tensor, k_in: torch.Tensor, v_in: torch.Tensor, nhead: int,
        key_padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Multi-head attention, batch-first: q_in (N,L,E), k_in/v_in (N,S,E) -> (N,L,E).

    Same arithmetic as ``nn.MultiheadAttention`` in eval mode (transformer.py:136,198-199;
    transformer_layers.py:51-52; tuber_ava.py:62) and as the reference's own
    ``MultiheadAttention.forward`` (transformer_layers.py:306-366): three slices of
    in_proj, q scaled by head_dim**-0.5 after its bias is added (:338), softmax over keys
    with padded keys at -inf, out_proj.  Dropout is the identity in eval mode.
    """
    e = q_in.shape[-1]
    hd = e // nhead
    w, b = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    q = F.linear(q_in, w[:e], b[:e]) * (float(hd) ** -0.5)
    k = F.linear(k_in, w[e:2 * e], b[e:2 * e])
    v = F.linear(v_in, w[2 * e:], b[2 * e:])
    n, l, _ = q.shape
    s = k.shape[1]
    q = q.view(n, l, nhead, hd).transpose(1, 2)
    k = k.view(n, s, nhead, hd).transpose(1, 2)
    v = v.view(n, s, nhead, hd).transpose(1, 2)
    scores = q @ k.transpose(-1, -2)                              # (N,H,L,S)
    if key_padding_mask is not None:
        scores = scores.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    out = torch.softmax(scores, dim=-1) @ v                       # (N,H,L,hd)
    out = out.transpose(1, 2).reshape(n, l, e)
    return F.linear(out, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def _ffn(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(F.relu(F.linear(x, sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"])),
                    sd[prefix + ".linear2.weight"], sd[prefix + ".linear2.bias"])


# ----------------------------------------------------------------------------------------
# backbone
# ----------------------------------------------------------------------------------------
def stage_strides(cfg) -> List[Tuple[int, int]]:
    """(temporal_stride, spatial_stride) per stage -- ir_CSN_152.py:124-135."""
    last = 2 if _m(cfg).LAST_STRIDE else 1
    return [(1, 1), (2, 2), (2, 2), (2, last)]


def bottleneck(sd, p: str, x: torch.Tensor, t_stride: int, s_stride: int, has_ds: bool) -> torch.Tensor:
    """ResNeXtBottleneck.forward (ir_CSN_152.py:70-90): 1x1x1 -> depthwise 3x3x3 (the stride
    lives here, :48-51) -> 1x1x1, BN after each, projection shortcut on a stage's first block."""
    planes = sd[p + ".conv1.weight"].shape[0]
    out = F.relu(_bn(sd, p + ".bn1", F.conv3d(x, sd[p + ".conv1.weight"])))
    out = F.conv3d(out, sd[p + ".conv3.weight"], stride=(t_stride, s_stride, s_stride), padding=1, groups=planes)
    out = F.relu(_bn(sd, p + ".bn3", out))
    out = _bn(sd, p + ".bn4", F.conv3d(out, sd[p + ".conv4.weight"]))
    if has_ds:   # ir_CSN_152.py:155-161: strided 1x1x1 conv + BN
        res = _bn(sd, p + ".down_sample.1",
                  F.conv3d(x, sd[p + ".down_sample.0.weight"], stride=(t_stride, s_stride, s_stride)))
    else:
        res = x
    return F.relu(out + res)


def csn_body(cfg, sd, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """ResNeXt.forward (ir_CSN_152.py:172-186): stem conv+BN+ReLU+maxpool, then 4 stages."""
    b = "backbone.body"
    x = F.conv3d(x, sd[b + ".conv1.weight"], stride=(1, 2, 2), padding=(1, 3, 3))
    x = F.relu(_bn(sd, b + ".bn1", x))
    x = F.max_pool3d(x, kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
    if taps is not None:
        taps["stem"] = x
    for li, ((ts, ss), nblk) in enumerate(zip(stage_strides(cfg), STAGE_BLOCKS[backbone_name(cfg)])):
        for bi in range(nblk):
            first = bi == 0
            x = bottleneck(sd, f"{b}.layer{li + 1}.{bi}", x, ts if first else 1, ss if first else 1, first)
        if taps is not None:
            taps[f"layer{li + 1}"] = x
    return x


def pool_decode(sd, xs: torch.Tensor) -> torch.Tensor:
    """'decode' temporal pool (backbone_builder.py:75-78; LSTRTransformerDecoder(Layer),
    transformer_layers.py:380-448): per pixel, one learned query token attends over the T'
    frame tokens; d=2048, 8 heads, no positional term, then the decoder's final LayerNorm."""
    bs, ch, t, h, w = xs.shape
    mem = xs.reshape(bs, ch, t, h * w).permute(0, 3, 2, 1).reshape(bs * h * w, t, ch)   # (pixels, T', C)
    tgt = sd["backbone.query_pool.weight"].view(1, 1, ch).expand(bs * h * w, 1, ch)
    p = "backbone.pool_decoder.layers.0"
    tgt = _ln(sd, p + ".norm1", tgt + mha(sd, p + ".self_attn", tgt, tgt, tgt, 8))
    tgt = _ln(sd, p + ".norm2", tgt + mha(sd, p + ".multihead_attn", tgt, mem, mem, 8))
    tgt = _ln(sd, p + ".norm3", tgt + _ffn(sd, p, tgt))
    tgt = _ln(sd, "backbone.pool_decoder.norm", tgt)                                  # (pixels,1,C)
    return tgt.view(bs, h * w, ch).permute(0, 2, 1).reshape(bs, ch, 1, h, w)


def temporal_pool(cfg, sd, xs: torch.Tensor) -> torch.Tensor:
    """backbone_builder.py:70-80."""
    m = _m(cfg)
    if not m.SINGLE_FRAME:
        return xs
    k = m.TEMP_LEN // m.DS_RATE
    if m.TEMPORAL_DS_STRATEGY == "avg":
        return F.avg_pool3d(xs, (k, 1, 1))
    if m.TEMPORAL_DS_STRATEGY == "max":
        return F.max_pool3d(xs, (k, 1, 1))
    if m.TEMPORAL_DS_STRATEGY == "decode":
        return pool_decode(sd, xs)
    t = xs.shape[2]
    return xs[:, :, t // 2: t // 2 + 1]


def resize_mask(mask: torch.Tensor, t: int, h: int, w: int) -> torch.Tensor:
    """(B,H,W) bool -> (B,t,h,w): nearest-neighbour resize then repeat over frames
    (backbone_builder.py:85-86).  Nearest picks source index floor(dst * in/out)."""
    b, hi, wi = mask.shape
    ys = torch.floor(torch.arange(h, dtype=torch.float32) * (hi / h)).long().clamp_(max=hi - 1)
    xs = torch.floor(torch.arange(w, dtype=torch.float32) * (wi / w)).long().clamp_(max=wi - 1)
    m = mask[:, ys][:, :, xs]
    return m[:, None].expand(b, t, h, w).contiguous()


def position_sine_3d(mask: torch.Tensor, d_model: int) -> torch.Tensor:
    """PositionEmbeddingSine_3D(normalize=True).forward (position_encoding.py:32-72).
    mask (B,T,H,W) bool -> (B, d_model, T, H, W): running counts of un-padded cells along
    t / y / x, normalised to 2*pi, divided by 10000**(2*(i//2)/n) with n = d/4 for t and 3d/8
    for y and x (:22-23), sin on even i and cos on odd i, concatenated (t, y, x)."""
    nt, ns = int(d_model / 8 * 2), int(d_model / 8 * 3)
    nm = (~mask).to(torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    te, ye, xe = nm.cumsum(1), nm.cumsum(2), nm.cumsum(3)
    te = te / (te[:, -1:] + eps) * scale
    ye = ye / (ye[:, :, -1:] + eps) * scale
    xe = xe / (xe[:, :, :, -1:] + eps) * scale

    def enc(e: torch.Tensor, n: int) -> torch.Tensor:
        i = torch.arange(n, dtype=torch.float32)
        div = 10000.0 ** (2 * torch.div(i, 2, rounding_mode="floor") / n)
        ph = e[..., None] / div
        out = torch.empty_like(ph)
        out[..., 0::2] = ph[..., 0::2].sin()
        out[..., 1::2] = ph[..., 1::2].cos()
        return out

    pos = torch.cat((enc(te, nt), enc(ye, ns), enc(xe, ns)), dim=-1)   # (B,T,H,W,d)
    return pos.permute(0, 4, 1, 2, 3).contiguous()


# ----------------------------------------------------------------------------------------
# DETR encoder / decoder (batch-first restatement of transformer.py)
# ----------------------------------------------------------------------------------------
def detr_encoder(cfg, sd, src: torch.Tensor, pos: torch.Tensor, kpm: torch.Tensor) -> torch.Tensor:
    """TransformerEncoderLayer.forward_post x ENC_LAYERS (transformer.py:153-168)."""
    m = _m(cfg)
    for i in range(m.ENC_LAYERS):
        p = f"transformer.encoder.layers.{i}"
        qk = src + pos
        src = _ln(sd, p + ".norm1", src + mha(sd, p + ".self_attn", qk, qk, src, m.NHEAD, kpm))
        src = _ln(sd, p + ".norm2", src + _ffn(sd, p, src))
    return src


def detr_decoder(cfg, sd, memory: torch.Tensor, pos: torch.Tensor, kpm: torch.Tensor) -> torch.Tensor:
    """TransformerDecoder with return_intermediate (transformer.py:99-128) over
    TransformerDecoderLayer.forward_post (:218-249).  Returns (L, B, Q, d)."""
    m = _m(cfg)
    bs = memory.shape[0]
    qpos = sd["query_embed.weight"][None].expand(bs, -1, -1)
    tgt = torch.zeros_like(qpos)                                   # transformer.py:60
    mem_k = memory + pos
    outs = []
    for i in range(m.DEC_LAYERS):
        p = f"transformer.decoder.layers.{i}"
        qk = tgt + qpos
        tgt = _ln(sd, p + ".norm1", tgt + mha(sd, p + ".self_attn", qk, qk, tgt, m.NHEAD))
        tgt = _ln(sd, p + ".norm2", tgt + mha(sd, p + ".multihead_attn", tgt + qpos, mem_k, memory, m.NHEAD, kpm))
        tgt = _ln(sd, p + ".norm3", tgt + _ffn(sd, p, tgt))
        outs.append(_ln(sd, "transformer.decoder.norm", tgt))
    return torch.stack(outs)


def class_encoder(sd, src_c: torch.Tensor, nhead: int = 8) -> torch.Tensor:
    """Class-branch TransformerEncoderLayer.forward_post (transformer_layers.py:71-97).
    src_c (B, d, T', H', W').  '_t' attends within a frame over H'W' positions, '_s' attends
    within a pixel over the T' frames (the reference's names are swapped); no positional
    term, no padding mask; the FFN sees the two normalised results concatenated and its
    residual is the un-attended input.  Returns tokens (B, T'*H'*W', d) in (t, y, x) order.

    The reference feeds DEC_LAYERS identical replicas of src_c through this layer
    (tuber_ava.py:133-135); the replicas never interact, so one pass per clip is the same
    function and that is what is evaluated here.
    """
    bs, d, t, h, w = src_c.shape
    hw = h * w
    x = src_c.reshape(bs, d, t, hw).permute(0, 2, 3, 1)            # (B,T',HW,d)
    p = "encoder.layers.0"
    xs = x.reshape(bs * t, hw, d)                                  # per-frame sequences
    a_t = _ln(sd, p + ".norm1_t", xs + mha(sd, p + ".self_attn_t", xs, xs, xs, nhead)).view(bs, t, hw, d)
    xt = x.permute(0, 2, 1, 3).reshape(bs * hw, t, d)              # per-pixel sequences
    a_s = _ln(sd, p + ".norm1_s", xt + mha(sd, p + ".self_attn_s", xt, xt, xt, nhead))
    a_s = a_s.view(bs, hw, t, d).permute(0, 2, 1, 3)
    cat = torch.cat((a_t, a_s), dim=-1)                            # (B,T',HW,2d)
    y = _ln(sd, p + ".norm2", x + _ffn(sd, p, cat))
    return y.reshape(bs, t * hw, d)


# ----------------------------------------------------------------------------------------
# full forward
# ----------------------------------------------------------------------------------------
@torch.no_grad()
def forward(cfg, sd: Dict[str, torch.Tensor], clips: torch.Tensor, mask: Optional[torch.Tensor] = None,
            taps: Optional[dict] = None, bank: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """DETR.forward (tuber_ava.py:97-148) in eval mode.

    clips (B,3,T,H,W) fp32, mask (B,H,W) bool (True = padding; None = no padding).
    Returns 'pred_logits' (L,B,Q,C), 'pred_boxes' (L,B,Q,4), 'pred_logits_b' (L,B,Q,3)
    [ava] or (L,B,2) [otherwise] for ALL decoder layers L (the reference's dict holds index
    -1 and, under AUX_LOSS, the first L-1 as 'aux_outputs', :144-157).

    ``bank`` (Bb, Nb, d) with Bb in {1, B}: long-term context window -- NO REFERENCE COUNTERPART ("parity unpinned"): the
    reference's README announces TubeR with long-term context but the code was never released.  Defined here after the paper:
    a clip's bank entry is class_proj(xt) averaged over the T' feature frames (returned as taps['bank_new'], (B, H'W', d)); the
    class-branch tokens attend over the window, post-norm: mem_c <- LN(mem_c + MHA(mem_c, bank, bank)) with ``ltc_attn`` / ``ltc_norm``.
    """
    m = _m(cfg)
    if m.NORMALIZE_BEFORE:
        raise NotImplementedError("pre-norm is broken in the reference (transformer.py:81,182)")
    bs = clips.shape[0]
    if mask is None:
        mask = torch.zeros((bs, clips.shape[3], clips.shape[4]), dtype=torch.bool)
    d = m.D_MODEL
    # -- backbone (backbone_builder.py:59-90)
    xt = csn_body(cfg, sd, clips, taps)                            # (B,2048,T',H',W')
    xs = temporal_pool(cfg, sd, xt)                                # (B,2048,T'',H',W')
    _, _, tp, hp, wp = xs.shape
    fmask = resize_mask(mask, tp, hp, wp)                          # (B,T'',H',W')
    pos = position_sine_3d(fmask, d)                               # (B,d,T'',H',W')
    if taps is not None:
        taps["xt"], taps["xs"], taps["pos"] = xt, xs, pos
    # -- DETR (tuber_ava.py:119; transformer.py:49-64)
    src = F.conv3d(xs, sd["input_proj.weight"], sd["input_proj.bias"])
    src = src.flatten(2).transpose(1, 2)                           # (B, T''H'W', d)
    pos_tok = pos.flatten(2).transpose(1, 2)
    kpm = fmask.flatten(1)
    memory = detr_encoder(cfg, sd, src, pos_tok, kpm)
    hs = detr_decoder(cfg, sd, memory, pos_tok, kpm)               # (L,B,Q,d)
    nl, _, nq, _ = hs.shape
    if taps is not None:
        taps["memory"], taps["hs"] = memory, hs
    # -- binary / actor-ness head (tuber_ava.py:121-125)
    if is_ava(cfg):
        logits_b = F.linear(hs, sd["class_embed_b.weight"], sd["class_embed_b.bias"])
    else:
        g = xt.mean(dim=(2, 3, 4))
        logits_b = F.linear(g, sd["class_embed_b.weight"], sd["class_embed_b.bias"])[None].expand(nl, -1, -1)
    # -- class branch (tuber_ava.py:129-141)
    src_c = F.conv3d(xt, sd["class_proj.weight"], sd["class_proj.bias"])
    mem_c = class_encoder(sd, src_c)                               # (B, T'H'W', d)
    if taps is not None:
        taps["mem_c"] = mem_c
        taps["bank_new"] = src_c.mean(dim=2).flatten(2).transpose(1, 2).contiguous()        # (B, H'W', d)
    if bank is not None:                                           # long-term context layer (in-repo definition, see docstring)
        kb = bank.expand(bs, -1, -1)
        mem_c = _ln(sd, "ltc_norm", mem_c + mha(sd, "ltc_attn", mem_c, kb, kb, 8))
        if taps is not None:
            taps["mem_ltc"] = mem_c
    q = hs.permute(1, 0, 2, 3).reshape(bs, nl * nq, d)             # every layer's queries of a clip
    q_class = mha(sd, "cross_attn", q, mem_c, mem_c, 8)
    q_class = q_class.view(bs, nl, nq, d).permute(1, 0, 2, 3)
    logits = F.linear(q_class, sd["class_fc.weight"], sd["class_fc.bias"])     # Dropout(0.5) = id in eval
    # -- box head (tuber_ava.py:142; criterion.py:494-497)
    x = hs
    for i in range(3):
        x = F.linear(x, sd[f"bbox_embed.layers.{i}.weight"], sd[f"bbox_embed.layers.{i}.bias"])
        if i < 2:
            x = F.relu(x)
    boxes = x.sigmoid()
    return {"pred_logits": logits.contiguous(), "pred_boxes": boxes.contiguous(),
            "pred_logits_b": logits_b.contiguous()}


def as_reference_dict(out: Dict[str, torch.Tensor], aux_loss: bool = True) -> dict:
    """Shape the all-layer outputs like the reference's return value (tuber_ava.py:144-157)."""
    res = {k: v[-1] for k, v in out.items()}
    if aux_loss:
        n = out["pred_logits"].shape[0]
        res["aux_outputs"] = [{k: v[i] for k, v in out.items()} for i in range(n - 1)]
    return res


# ----------------------------------------------------------------------------------------
# post-processing (SURVEY section 8f "next" row 1)
# ----------------------------------------------------------------------------------------
def postprocess_ava(logits: torch.Tensor, boxes: torch.Tensor, logits_b: torch.Tensor,
                    target_sizes: torch.Tensor):
    """PostProcessAVA.forward (criterion.py:447-482): scores = sigmoid(logits) * p_actor gated at
    0.8, boxes cxcywh -> xyxy scaled to (w,h,w,h)."""
    pb = logits_b.softmax(-1)[..., 1:2]
    scores = logits.sigmoid() * ((pb > 0.8).float() * pb)
    cx, cy, w, h = boxes.unbind(-1)
    xyxy = torch.stack((cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h), dim=-1)
    img_h, img_w = target_sizes.unbind(1)
    xyxy = xyxy * torch.stack((img_w, img_h, img_w, img_h), dim=1)[:, None, :]
    return scores, xyxy, pb


def rel_err(a: torch.Tensor, b: torch.Tensor) -> Tuple[float, float]:
    """(max|a-b| / max|b|, ||a-b||2 / ||b||2) -- the parity metric of SURVEY section 8d."""
    a = a.double().flatten()
    b = b.double().flatten()
    return (float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)),
            float((a - b).norm() / b.norm().clamp_min(1e-30)))
