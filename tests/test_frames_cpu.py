"""CPU: the frame-loading oracle (oracle/frame_oracle.py: baseline JPEG decode as libjpeg does it + Pillow's bicubic resize) against
pixels produced by Pillow itself (tests/golden/frames.npz, oracle/make_golden_frames.py) -- bit for bit."""
import os

import numpy as np
import pytest

from oracle import frame_oracle as F

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames.npz"))
JPEG_CASES = sorted({k.split("/")[0] for k in GOLD.files if k.endswith("/jpeg")})
RESIZE_CASES = sorted({k.split("/")[0] for k in GOLD.files if k.endswith("/src")})


@pytest.mark.parametrize("name", JPEG_CASES)
def test_decode_matches_pillow(name):
    px = F.decode_jpeg(GOLD[name + "/jpeg"].tobytes())
    assert px.dtype == np.uint8 and np.array_equal(px, GOLD[name + "/pixels"])


@pytest.mark.parametrize("name", JPEG_CASES)
def test_decode_then_resize_matches_pillow(name):
    want = GOLD[name + "/resized"]
    got = F.load_frame(GOLD[name + "/jpeg"].tobytes(), want.shape[0], want.shape[1])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", RESIZE_CASES)
def test_resize_matches_pillow(name):
    want = GOLD[name + "/dst"]
    assert np.array_equal(F.resize_bicubic(GOLD[name + "/src"], want.shape[0], want.shape[1]), want)


def test_live_pillow_if_present():
    """the same comparison against the Pillow of the machine the tests run on (skipped without Pillow)"""
    PIL = pytest.importorskip("PIL")
    import io
    from PIL import Image
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, size=(45, 61, 3), dtype=np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=88)
    ref = Image.open(io.BytesIO(buf.getvalue()))
    assert np.array_equal(F.decode_jpeg(buf.getvalue()), np.asarray(ref))
    assert np.array_equal(F.load_frame(buf.getvalue(), 30, 50), np.asarray(ref.resize((50, 30))))


def test_rejects_what_it_does_not_restate():
    with pytest.raises(AssertionError):
        F.parse_jpeg(b"not a jpeg")
    PIL = pytest.importorskip("PIL")
    import io
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(np.zeros((16, 16, 3), dtype=np.uint8)).save(buf, format="JPEG", progressive=True)
    with pytest.raises(ValueError):
        F.decode_jpeg(buf.getvalue())


@pytest.mark.parametrize("name", JPEG_CASES)
def test_host_entropy_decoder_matches_the_oracle(name):
    """the product's host-side half (marker parsing + Huffman decoding in csrc/frames.cu, tuber_op_jpeg_coefficients: no device needed)
    against the oracle's coefficient arrays; truncated files are refused, a missing EOI marker is tolerated"""
    import ctypes as C
    import tuber_b200  # noqa: F401
    from tuber_b200 import _lib
    lib = _lib.load()
    data = GOLD[name + "/jpeg"].tobytes()
    info = (C.c_int32 * 10)()
    assert lib.tuber_op_jpeg_coefficients(data, len(data), None, 0, info) == 0
    _, coefs = F.decode_coefficients(data)
    assert [info[4 + 2 * c] * info[5 + 2 * c] for c in range(3)] == [c.shape[0] * c.shape[1] for c in coefs]
    total = sum(c.size for c in coefs)
    buf = np.zeros(total, dtype=np.int16)
    assert lib.tuber_op_jpeg_coefficients(data, len(data), buf.ctypes.data_as(C.c_void_p), total, info) == 0
    assert np.array_equal(buf, np.concatenate([c.reshape(-1) for c in coefs]).astype(np.int16))
    assert lib.tuber_op_jpeg_coefficients(data, len(data), buf.ctypes.data_as(C.c_void_p), total - 1, info) != 0      # buffer too small
    assert lib.tuber_op_jpeg_coefficients(data[:-2], len(data) - 2, buf.ctypes.data_as(C.c_void_p), total, info) == 0  # no EOI
    half = data[: len(data) // 2]
    assert lib.tuber_op_jpeg_coefficients(half, len(half), buf.ctypes.data_as(C.c_void_p), total, info) != 0
    assert b"truncated" in lib.tuber_frames_last_error()
