"""GPU JPEG codec (libr360codec.so over nvJPEG): the library exports what include/remap360_codec.h declares
(CPU), and decode / encode round trips agree with OpenCV's codec to JPEG accuracy (``-m gpu``)."""

import pathlib
import re
import subprocess

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_codec_library_exports_exactly_the_declared_symbols():
    pytest.importorskip("torch")
    from remap360 import codec
    header = (ROOT / "include" / "remap360_codec.h").read_text()
    declared = set(re.findall(r"^(?:int|void|const char\*)\s+(r360_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(codec.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", str(codec.CODEC_PATH)], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T r360_" in line}
    assert exported == declared
    lib = codec.load()
    for name in declared:
        assert hasattr(lib, name)


def _smooth(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([0.5 + 0.3 * np.sin(xx / w * 9 + k) * np.cos(yy / h * 7 - k) + 0.1 * np.sin((xx + 2 * yy) / 40 + k)
                    for k in range(3)], axis=-1)
    return np.clip(np.rint(img * 255 + rng.normal(0, 2, img.shape)), 0, 255).astype(np.uint8)


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


@pytest.mark.gpu
def test_decode_matches_opencv_and_encode_round_trips():
    torch = pytest.importorskip("torch")
    cv2 = pytest.importorskip("cv2")
    from remap360 import codec
    rng = np.random.default_rng(3)
    jc = codec.JpegCodec()
    for (h, w) in ((240, 320), (517, 771)):
        img = _smooth(rng, h, w)                                       # B, G, R as cv2 holds it
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, 0x111111])
        data = enc.tobytes()
        assert jc.info(data) == (w, h, 3)
        got = jc.decode(data).cpu().numpy()
        want = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)
        assert got.shape == want.shape and _psnr(got, want) > 45          # same stream, two IDCT implementations
        assert _psnr(jc.decode(data, channel_order="rgb").cpu().numpy()[..., ::-1], want) > 45
        # 4:2:0 input (what cameras write)
        ok, enc420 = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
        assert _psnr(jc.decode(enc420.tobytes()).cpu().numpy(), cv2.imdecode(enc420, cv2.IMREAD_UNCHANGED)) > 38
        # encode on the GPU, decode with OpenCV
        dev = torch.from_numpy(img).cuda()
        for q, floor in ((95, 40.0), (100, 44.0), (60, 30.0)):
            out = jc.encode(dev, quality=q)
            assert out[:2] == b"\xff\xd8" and out[-2:] == b"\xff\xd9"
            back = cv2.imdecode(np.frombuffer(out, np.uint8), cv2.IMREAD_UNCHANGED)
            assert back.shape == img.shape and _psnr(back, img) > floor, (q, _psnr(back, img))
        # row-padded views (alloc_views) encode without a copy
        import remap360
        padded = remap360.alloc_views(1, 1, h, w, 3, torch.uint8, "cuda")[0, 0]
        padded.copy_(dev)
        back = cv2.imdecode(np.frombuffer(jc.encode(padded, 95), np.uint8), cv2.IMREAD_UNCHANGED)
        assert _psnr(back, img) > 40
    grey = _smooth(rng, 200, 264)[..., :1].copy()
    back = cv2.imdecode(np.frombuffer(jc.encode(torch.from_numpy(grey).cuda(), 95), np.uint8), cv2.IMREAD_UNCHANGED)
    assert back.shape == grey.shape[:2] and _psnr(back, grey[..., 0]) > 40
    with pytest.raises(codec.CodecError):
        jc.decode(b"not a jpeg at all")
