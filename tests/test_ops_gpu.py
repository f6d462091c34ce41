"""GPU: every single-operator entry point of the C-ABI against a torch fp64 restatement of the op."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def abi():
    from tests import _abi
    return _abi


def test_split_roundtrip(abi):
    x = torch.randn(257, 192, device="cuda") * 3
    y = abi.from_split(abi.to_split(x), 257, 192)
    assert float(((y - x).abs() / x.abs().clamp_min(1e-6)).max()) < 2.0 ** -15


@pytest.mark.parametrize("m,n,k", [(300, 64, 64), (128, 128, 256), (1000, 256, 2048), (4096, 512, 128), (77, 192, 512),
                                   (20000, 256, 64), (16384, 256, 1024), (13000, 512, 512), (16484, 256, 1024)])
@pytest.mark.parametrize("variant", ["plain", "bn_relu_res", "split_out_split_res", "res_mod", "mixed_res_f32_out_split", "split_out"])
def test_gemm_tc(abi, m, n, k, variant):
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / math.sqrt(k)
    scale = shift = res = None
    kw = {}
    ref = a.double() @ w.double().t()
    if variant != "plain":
        scale = torch.rand(n, device="cuda", generator=g) + 0.5
        shift = torch.randn(n, device="cuda", generator=g)
        ref = ref * scale.double() + shift.double()
    if variant in ("bn_relu_res", "split_out_split_res"):
        res = torch.randn(m, n, device="cuda", generator=g)
        ref = torch.relu(ref + res.double())
        kw = dict(relu=True, res_split=variant == "split_out_split_res", c_split=variant == "split_out_split_res")
    if variant == "mixed_res_f32_out_split":
        res = torch.randn(m, n, device="cuda", generator=g)
        ref = ref + res.double()
        kw = dict(c_split=True)
    if variant == "split_out":
        ref = torch.relu(ref)
        kw = dict(relu=True, c_split=True)
    if variant == "res_mod":
        mod = 13
        res = torch.randn(mod, n, device="cuda", generator=g)
        ref = ref + res.double()[torch.arange(m, device="cuda") % mod]
        kw = dict(res_mod=mod)
    out = abi.gemm_tc(a, w, scale, shift, res, **kw)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 3e-5, (m, n, k, variant)


@pytest.mark.parametrize("m,n,k", [(16384, 1024, 256), (9500, 1024, 128), (20000, 512, 256)])
@pytest.mark.parametrize("variant", ["plain", "bn_relu_res", "split_out_split_res", "res_mod", "mixed_res_f32_out_split", "split_out"])
def test_gemm_tc_pair128(abi, monkeypatch, m, n, k, variant):
    """the same problems through gemm2_bf16x3_kernel<3, 2, 2, 128> (CTA pairs, 256 x 128 tiles; TUBER_PAIR128 is read per call)"""
    monkeypatch.setenv("TUBER_PAIR128", "1")
    test_gemm_tc(abi, m, n, k, variant)


@pytest.mark.parametrize("m,k,kb,n1,n2,with_res", [(1000, 64, 0, 256, 64, True), (128, 64, 64, 256, 64, False), (40000, 64, 0, 256, 64, True),
                                                    (5000, 64, 0, 256, 128, True), (33333, 64, 64, 256, 128, False),
                                                    (300, 128, 0, 256, 64, True), (50000, 64, 64, 256, 64, False),
                                                    (70001, 128, 0, 256, 128, True), (700, 128, 0, 512, 128, True),
                                                    (45000, 128, 0, 512, 128, True), (30001, 128, 256, 512, 128, False),
                                                    (20000, 128, 0, 512, 256, True), (5001, 128, 256, 512, 256, False),
                                                    (16384, 256, 0, 1024, 256, True), (900, 256, 512, 1024, 256, False),
                                                    (40001, 256, 0, 1024, 256, True)])
def test_gemm_tc_fused2(abi, m, k, kb, n1, n2, with_res):
    """conv4 (+ residual / K-concatenated shortcut) chained with the next block's conv1 through the shared-memory panels"""
    g = torch.Generator(device="cuda").manual_seed(m + k + kb + n2)
    a = torch.randn(m, k, device="cuda", generator=g)
    ab = torch.randn(m, kb, device="cuda", generator=g) if kb else None
    w = torch.randn(n1, k + kb, device="cuda", generator=g) / math.sqrt(k + kb)
    scale, shift = torch.rand(n1, device="cuda", generator=g) + 0.5, torch.randn(n1, device="cuda", generator=g) * 0.1
    res = torch.randn(m, n1, device="cuda", generator=g) if with_res else None
    w2 = torch.randn(n2, n1, device="cuda", generator=g) / math.sqrt(n1)
    scale2, shift2 = torch.rand(n2, device="cuda", generator=g) + 0.5, torch.randn(n2, device="cuda", generator=g) * 0.1
    x, t1 = abi.gemm_tc_fused2(a, w, scale, shift, res, w2, scale2, shift2, ab)
    acat = a.double() if ab is None else torch.cat([a, ab], 1).double()
    ref_x = acat @ w.double().t() * scale.double() + shift.double()
    if res is not None:
        ref_x = ref_x + res.double()
    ref_x = torch.relu(ref_x)
    ref_t = torch.relu(ref_x @ w2.double().t() * scale2.double() + shift2.double())
    assert _rel(x, ref_x) < 3e-5
    assert _rel(t1, ref_t) < 3e-5


@pytest.mark.parametrize("m,k,kb,with_res", [(700, 128, 0, True), (45000, 128, 0, True), (30001, 128, 256, False), (131072, 128, 0, True)])
def test_gemm_tc_fused2_pair_512(abi, monkeypatch, m, k, kb, with_res):
    """the 512 -> 128 shapes through the CTA-pair form gemm_fused2p_kernel<512, 128> (TUBER_FUSE2P_L2 is read per call)"""
    monkeypatch.setenv("TUBER_FUSE2P_L2", "1")
    test_gemm_tc_fused2(abi, m, k, kb, 512, 128, with_res)


@pytest.mark.parametrize("m,n,k,act", [(90, 3, 256, 0), (90, 4, 256, 2), (720, 80, 256, 0), (8, 2, 2048, 0), (333, 256, 64, 1),
                                       (15360, 22, 256, 0), (15361, 4, 256, 2), (2049, 3, 256, 1), (3, 5, 1024, 0)])
def test_sgemm(abi, m, n, k, act):
    a = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, device="cuda") / math.sqrt(k)
    b = torch.randn(n, device="cuda")
    ref = a.double() @ w.double().t() + b.double()
    ref = torch.relu(ref) if act == 1 else (torch.sigmoid(ref) if act == 2 else ref)
    assert _rel(abi.sgemm(a, w, b, None, act), ref) < 2e-6


@pytest.mark.parametrize("c,st_t,st_s,dims", [(64, 1, 1, (2, 4, 9, 10)), (128, 2, 2, (1, 8, 16, 16)), (512, 2, 1, (2, 3, 5, 7)),
                                              (256, 2, 2, (1, 5, 7, 9)), (64, 1, 1, (1, 8, 32, 32)), (128, 1, 1, (2, 5, 16, 17)),
                                              (512, 1, 1, (1, 1, 3, 3)), (256, 1, 1, (1, 2, 8, 8)), (128, 2, 2, (2, 32, 64, 64)),
                                              (64, 2, 2, (1, 7, 19, 35))])
def test_dwconv(abi, c, st_t, st_s, dims):
    from tuber_b200 import _lib
    b, t, h, w = dims
    x = torch.randn(b, c, t, h, w, device="cuda")
    wt = torch.randn(c, 1, 3, 3, 3, device="cuda") * 0.3
    scale, shift = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    ref = F.conv3d(x.double(), wt.double(), stride=(st_t, st_s, st_s), padding=1, groups=c)
    ref = torch.relu(ref * scale.double()[None, :, None, None, None] + shift.double()[None, :, None, None, None])
    xin = x.permute(0, 2, 3, 4, 1).contiguous()
    wpk = wt.reshape(c, 27).t().contiguous()
    to, ho, wo = ref.shape[2:]
    out = torch.empty((b * to * ho * wo, c), device="cuda")
    _lib.check(_lib.load().tuber_op_dwconv(abi.P(xin), abi.P(wpk), abi.P(scale), abi.P(shift), abi.P(out), b, t, h, w, c, st_t, st_s,
                                           abi.stream()))
    got = abi.from_split(out, b * to * ho * wo, c).view(b, to, ho, wo, c).permute(0, 4, 1, 2, 3)
    assert _rel(got, ref) < 3e-5


@pytest.mark.parametrize("dims", [(1, 4, 32, 32), (2, 3, 45, 70), (1, 2, 64, 300), (1, 3, 130, 256),
                                  # conv rows wider than 128 outputs: column tiles of the pair kernel with a one-column pool halo --
                                  # the evaluation widths 341 / 455 (W1 = 171 / 228), one column past a tile (257, 258), three tiles (520)
                                  (1, 2, 40, 341), (1, 2, 34, 455), (1, 1, 20, 257), (1, 1, 20, 258), (1, 1, 18, 520), (1, 1, 12, 506)])
def test_stem(abi, dims):
    from tuber_b200 import _lib
    b, t, h, w = dims
    x = torch.randn(b, 3, t, h, w, device="cuda")
    wt = torch.randn(64, 3, 3, 7, 7, device="cuda") * 0.05
    scale, shift = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.1
    ref = F.conv3d(x.double(), wt.double(), stride=(1, 2, 2), padding=(1, 3, 3))
    ref = torch.relu(ref * scale.double()[None, :, None, None, None] + shift.double()[None, :, None, None, None])
    refp = F.max_pool3d(ref, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    h1, w1, h2, w2 = ref.shape[3], ref.shape[4], refp.shape[3], refp.shape[4]
    wpk = wt.reshape(64, 441).contiguous()
    conv = torch.empty((b, t, h1, w1, 64), device="cuda")
    pooled = torch.empty((b * t * h2 * w2, 64), device="cuda")
    _lib.check(_lib.load().tuber_op_stem(abi.P(x), abi.P(wpk), abi.P(scale), abi.P(shift), abi.P(conv), abi.P(pooled), b, t, h, w,
                                         abi.stream()))
    assert _rel(conv.permute(0, 4, 1, 2, 3), ref) < 3e-5
    got = abi.from_split(pooled, b * t * h2 * w2, 64).view(b, t, h2, w2, 64).permute(0, 4, 1, 2, 3)
    assert _rel(got, refp) < 3e-5


@pytest.mark.parametrize("dims", [(1, 4, 32, 32), (2, 3, 45, 70), (1, 3, 130, 256), (2, 2, 256, 256), (1, 2, 40, 341), (1, 1, 20, 258),
                                  (1, 1, 18, 520), (3, 5, 70, 64)])
def test_stem_two_rows_per_batch(abi, monkeypatch, dims):
    """stem_tc3_kernel (TUBER_STEM3 is read per call): two conv rows per accumulator batch sharing their A K-steps"""
    monkeypatch.setenv("TUBER_STEM3", "1")
    test_stem(abi, dims)


@pytest.mark.parametrize("c", [256, 2048])
def test_layernorm(abi, c):
    from tuber_b200 import _lib
    x, r = torch.randn(77, c, device="cuda"), torch.randn(77, c, device="cuda")
    g, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    out = torch.empty_like(x)
    _lib.check(_lib.load().tuber_op_layernorm(abi.P(x), abi.P(r), abi.P(g), abi.P(b), abi.P(out), 77, c, abi.stream()))
    ref = F.layer_norm((x + r).double(), (c,), g.double(), b.double(), 1e-5)
    assert _rel(out, ref) < 5e-6


@pytest.mark.parametrize("nb,h,l,s,d,masked", [(2, 8, 256, 256, 32, True), (3, 8, 15, 15, 32, False), (2, 8, 15, 300, 32, True),
                                               (5, 8, 1, 4, 256, False), (2, 8, 90, 1024, 32, False), (4, 8, 4, 4, 32, False),
                                               (2, 8, 40, 33, 32, True), (1, 8, 320, 320, 32, False), (3, 8, 7, 17, 32, True),
                                               (2, 8, 64, 64, 32, False), (2, 8, 130, 70, 32, True), (3, 8, 8, 8, 32, True),
                                               (2, 16, 2, 2, 32, False), (2, 8, 16, 129, 32, False),
                                               # long sequences, no mask: the tcgen05 flash-attention kernel (attn_tc.cu) -- exact tiles,
                                               # query / key tails, one chunk pair, many chunks
                                               (1, 8, 128, 1024, 32, False), (2, 4, 300, 2500, 32, False), (1, 2, 129, 1025, 32, False),
                                               (1, 8, 256, 16384, 32, False),
                                               # the reference path's own attention sites on the tcgen05 kernel: encoder self-attention
                                               # (L = S = H'W' = 256; 352 at 256 x 341 input) with and without a key padding mask (one
                                               # sequence loses a whole chunk, one every third key), the class branch's spatial attention
                                               # (batch T'B), JHMDB's decoder self-attention (320) and class cross-attention (1920 x 512),
                                               # a ragged feature map (15 x 22 = 330 keys), the smallest supported tile
                                               (2, 8, 256, 256, 32, True), (2, 8, 256, 256, 32, False), (4, 8, 352, 352, 32, True),
                                               (32, 8, 256, 256, 32, False), (2, 8, 320, 320, 32, True), (2, 8, 1920, 512, 32, False),
                                               (2, 8, 200, 330, 32, True), (3, 8, 64, 128, 32, True)])
def test_attention(abi, nb, h, l, s, d, masked):
    from tuber_b200 import _lib
    kernel = _lib.load().tuber_op_attention_kernel(nb, h, l, s, d, int(masked)).decode()
    if d == 32 and l >= 64 and s >= 128:                      # the tcgen05 kernel takes every such shape, masked or not
        assert kernel == "attn_tc_kernel", kernel
    else:
        assert kernel != "attn_tc_kernel", kernel
    e = h * d
    g = torch.Generator(device="cuda").manual_seed(nb * 1000 + l * 7 + s)
    q, k, v = (torch.randn(nb, n, e, device="cuda", generator=g) for n in (l, s, s))
    kpm = None
    if masked:
        kpm = torch.zeros(nb, s, dtype=torch.uint8, device="cuda")
        kpm[0, s // 2:] = 1
        kpm[1, ::3] = 1
    out = torch.empty(nb, l, e, device="cuda")
    scale = d ** -0.5
    _lib.check(_lib.load().tuber_op_attention(abi.P(q), abi.P(k), abi.P(v), abi.P(kpm), abi.P(out), nb, h, l, s, d, C.c_float(scale),
                                              abi.stream()))
    qh, kh, vh = (t.double().view(nb, -1, h, d).transpose(1, 2) for t in (q, k, v))
    sc = qh @ kh.transpose(-1, -2) * scale
    if kpm is not None:
        sc = sc.masked_fill(kpm.bool()[:, None, None, :], float("-inf"))
    ref = (sc.softmax(-1) @ vh).transpose(1, 2).reshape(nb, l, e)
    assert _rel(out, ref) < 3e-5          # bf16x3 operand splitting (2^-16 per operand); the path's bar is 1e-3


@pytest.mark.parametrize("nb,h,l,s", [(2, 4, 300, 2500), (1, 8, 128, 1024), (3, 2, 129, 1153)])
def test_attention_tc_preconverted_keys(abi, nb, h, l, s, monkeypatch):
    """The tcgen05 attention with K / V converted once into per-chunk tile images (the form the plan uses, attn_tc.cu PREP) gives
    bit for bit what the same kernel gives when every CTA converts its own chunks."""
    from tuber_b200 import _lib
    d, e = 32, h * 32
    g = torch.Generator(device="cuda").manual_seed(s)
    q, k, v = (torch.randn(nb, n, e, device="cuda", generator=g) for n in (l, s, s))
    outs = []
    for prep in ("0", "1"):
        monkeypatch.setenv("TUBER_OP_ATTN_PREP", prep)
        out = torch.empty(nb, l, e, device="cuda")
        _lib.check(_lib.load().tuber_op_attention(abi.P(q), abi.P(k), abi.P(v), None, abi.P(out), nb, h, l, s, d, C.c_float(d ** -0.5),
                                                  abi.stream()))
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])


def test_posenc(abi):
    from oracle import tuber_oracle as O
    from tuber_b200 import _lib
    b, t, h, w = 3, 2, 6, 8
    mask = torch.zeros(b, t, h, w, dtype=torch.bool)
    mask[1, :, :, 5:] = True
    mask[2, :, 4:, :] = True
    ref = O.position_sine_3d(mask, 256).flatten(2).transpose(1, 2)           # (B, THW, 256)
    m8 = mask.to(torch.uint8).cuda()
    out = torch.empty(b, t * h * w, 256, device="cuda")
    _lib.check(_lib.load().tuber_op_posenc(abi.P(m8), abi.P(out), b, t, h, w, 256, abi.stream()))
    torch.cuda.synchronize()
    assert float((out.cpu() - ref).abs().max()) < 2e-5


@pytest.mark.parametrize("shape", [(2, 4, 12, 20), (1, 3, 7, 9), (3, 8, 64, 64)])
@pytest.mark.parametrize("norm", ["imagenet", "other"])
def test_normalize_u8(abi, shape, norm):
    """uint8 frames -> normalised fp32 clip: bit-identical to the reference's ToTensor + Normalize (restated by oracle.frames_to_clips,
    pinned by tests/golden/input_u8.npz); T*H*W % 4 != 0 exercises the scalar path."""
    from oracle import tuber_oracle as O
    from tuber_b200 import _lib
    mean, std = (O.IMAGENET_MEAN, O.IMAGENET_STD) if norm == "imagenet" else ((0.45, 0.5, 0.375), (0.225, 0.3, 0.25))
    fr = O.make_frames_u8(*shape, seed=11)
    ref = O.frames_to_clips(fr, mean, std)
    B, T, H, W = shape
    out = torch.empty((B, 3, T, H, W), device="cuda", dtype=torch.float32)
    m, s = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _lib.check(_lib.load().tuber_op_normalize_u8(abi.P(fr.cuda()), m, s, abi.P(out), B, T * H * W, abi.stream()))
    assert torch.equal(out.cpu(), ref)


def test_normalize_u8_golden(abi):
    import os
    import numpy as np
    from tuber_b200 import _lib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_u8.npz"))
    for name in ("a", "b", "table"):
        fr, ref = torch.from_numpy(g[f"frames_{name}"]), torch.from_numpy(g[f"clips2_{name}"])
        B, T, H, W, _ = fr.shape
        out = torch.empty((B, 3, T, H, W), device="cuda", dtype=torch.float32)
        m, s = (C.c_float * 3)(*[float(v) for v in g["mean2"]]), (C.c_float * 3)(*[float(v) for v in g["std2"]])
        _lib.check(_lib.load().tuber_op_normalize_u8(abi.P(fr.cuda()), m, s, abi.P(out), B, T * H * W, abi.stream()))
        assert torch.equal(out.cpu(), ref), name
