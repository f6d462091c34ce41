"""GPU: tuber_frames_decode (host Huffman decoding + CUDA inverse DCT / upsampling / colour conversion / resize, csrc/frames.cu)
against pixels produced by Pillow itself (tests/golden/frames.npz) and against the CPU oracle -- bit for bit -- and the whole
chain JPEG files -> detections against the fp32 entry point on the oracle's clip."""
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames.npz"))
JPEG_CASES = sorted({k.split("/")[0] for k in GOLD.files if k.endswith("/jpeg")})


@pytest.fixture(scope="module")
def dec():
    import tuber_b200
    d = tuber_b200.FrameDecoder()
    yield d
    d.close()


@pytest.mark.parametrize("name", JPEG_CASES)
def test_decode_and_resize_match_pillow(dec, name):
    data = GOLD[name + "/jpeg"].tobytes()
    px, rs = GOLD[name + "/pixels"], GOLD[name + "/resized"]
    got = dec.decode([data], px.shape[0], px.shape[1])              # same size: decode only
    assert np.array_equal(got[0].cpu().numpy(), px)
    got = dec.decode([data], rs.shape[0], rs.shape[1])              # the reference's Image.open + resize
    assert np.array_equal(got[0].cpu().numpy(), rs)


def test_mixed_batch_matches_the_oracle(dec):
    """frames of different sizes and chroma subsamplings in one call, one output size; many frames (more than host threads)"""
    from oracle import frame_oracle as F
    datas = [GOLD[n + "/jpeg"].tobytes() for n in JPEG_CASES] * 5
    got = dec.decode(datas, 40, 56).cpu().numpy()
    for i, d in enumerate(datas[:len(JPEG_CASES)]):
        want = F.load_frame(d, 40, 56)
        assert np.array_equal(got[i], want), JPEG_CASES[i]
        assert np.array_equal(got[i + 3 * len(JPEG_CASES)], want)
    out = torch.empty((2, 24, 24, 3), dtype=torch.uint8, device="cuda")
    assert dec.decode(datas[:2], 24, 24, out=out) is out


def test_real_frame_size_against_live_pillow(dec):
    """a 360 x 480 frame resized to the AVA evaluation size 256 x 341, and one axis unchanged, against the Pillow of this machine"""
    Image = pytest.importorskip("PIL.Image")
    from oracle.make_golden_frames import synth
    import tuber_b200
    datas, want = [], []
    for i in range(3):
        buf = io.BytesIO()
        Image.fromarray(synth(360, 480, 40 + i)).save(buf, format="JPEG", quality=70 + 10 * i)
        datas.append(buf.getvalue())
    h, w = tuber_b200.clip_size(360, 480, 256)
    assert (h, w) == (256, 341)
    got = dec.decode(datas, h, w).cpu().numpy()
    for i, d in enumerate(datas):
        assert np.array_equal(got[i], np.asarray(Image.open(io.BytesIO(d)).resize((w, h))))
    got = dec.decode(datas, 360, 300).cpu().numpy()                 # horizontal pass only
    assert np.array_equal(got[1], np.asarray(Image.open(io.BytesIO(datas[1])).resize((300, 360))))
    got = dec.decode(datas, 200, 480).cpu().numpy()                 # vertical pass only
    assert np.array_equal(got[2], np.asarray(Image.open(io.BytesIO(datas[2])).resize((480, 200))))


def test_unsupported_input_is_refused(dec):
    import tuber_b200
    with pytest.raises(tuber_b200.frames.FrameDecodeError):
        dec.decode([b"not a jpeg at all"], 8, 8)
    Image = pytest.importorskip("PIL.Image")
    buf = io.BytesIO()
    Image.fromarray(np.zeros((16, 16, 3), dtype=np.uint8)).save(buf, format="JPEG", progressive=True)
    with pytest.raises(tuber_b200.frames.FrameDecodeError, match="baseline"):
        dec.decode([buf.getvalue()], 16, 16)
    buf = io.BytesIO()
    Image.fromarray(np.zeros((16, 16), dtype=np.uint8)).save(buf, format="JPEG")           # greyscale
    with pytest.raises(tuber_b200.frames.FrameDecodeError):
        dec.decode([buf.getvalue()], 16, 16)
    good = GOLD[JPEG_CASES[0] + "/jpeg"].tobytes()
    with pytest.raises(tuber_b200.frames.FrameDecodeError):
        dec.decode([good[: len(good) // 2]], 16, 16)                                       # truncated scan
    with pytest.raises(ValueError):
        dec.decode([], 16, 16)


def test_jpeg_files_to_detections(dec, tmp_path):
    """the widened path end to end: JPEG files -> FrameDecoder -> forward_raw_u8, against forward_raw on the clip the reference's
    loader + transform (Pillow decode / resize, ToTensor, Normalize: oracle.frames_to_clips) would have produced -- bit for bit"""
    Image = pytest.importorskip("PIL.Image")
    import tuber_b200
    from oracle import tuber_oracle as O
    from oracle.cases import build_case
    from oracle.make_golden_frames import synth
    cfg, sd, clips, _ = build_case("A_csn50_avg_bnrand")
    B, _, T, H, W = clips.shape
    model, _, _ = tuber_b200.build_model(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    paths = []
    for i in range(B * T):
        p = tmp_path / f"img_{i:05d}.jpg"
        Image.fromarray(synth(90, 120, 300 + i)).save(p, format="JPEG", quality=80)
        paths.append(str(p))
    frames = dec.load_clip(paths, H, W).view(B, T, H, W, 3)
    ref_frames = torch.from_numpy(np.stack([np.asarray(Image.open(p).resize((W, H))) for p in paths])).view(B, T, H, W, 3)
    assert torch.equal(frames.cpu(), ref_frames)
    got = model.forward_raw_u8(frames)
    want = model.forward_raw(O.frames_to_clips(ref_frames).cuda())
    for k in want:
        assert torch.equal(got[k], want[k]), k
