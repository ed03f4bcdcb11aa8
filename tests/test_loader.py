"""Data path in front of the GPU preprocessing: .flo reader on the CPU, the pinned double-buffered
loader on the GPU against the reference loaders' arithmetic (oracle.preprocess_frame / preprocess_flow on the same decodes)."""
import os

import numpy as np
import pytest
import torch

import ammc_oracle as O
import ammcnet_aaai2021_b200 as A


def _write_video(root, name, T, h0, w0, seed):
    import cv2
    rng = np.random.default_rng(seed)
    rd, od = os.path.join(root, "rgb", name), os.path.join(root, "op", name)
    os.makedirs(rd), os.makedirs(od)
    base = rng.integers(0, 256, (h0 // 8 + 1, w0 // 8 + 1, 3), dtype=np.uint8)
    for t in range(T):
        img = cv2.resize(np.roll(base, t, axis=1), (w0, h0), interpolation=cv2.INTER_CUBIC)      # smooth: JPEG-friendly
        cv2.imwrite(os.path.join(rd, "%04d.jpg" % t), img, [cv2.IMWRITE_JPEG_QUALITY, 95])
    for t in range(T - 1):
        A.write_flo(os.path.join(od, "%04d.flo" % t), rng.standard_normal((h0, w0, 2)).astype(np.float32))
    return rd, od


def test_flo_roundtrip_and_bad_files(tmp_path):
    flow = np.random.default_rng(1).standard_normal((7, 11, 2)).astype(np.float32)
    p = str(tmp_path / "a.flo")
    A.write_flo(p, flow)
    assert np.array_equal(A.read_flo(p), flow)
    with open(str(tmp_path / "bad.flo"), "wb") as f:
        f.write(b"\x00" * 64)
    with pytest.raises(ValueError, match="not a .flo"):
        A.read_flo(str(tmp_path / "bad.flo"))
    with open(p, "r+b") as f:
        f.truncate(100)
    with pytest.raises(ValueError, match="truncated"):
        A.read_flo(p)


@pytest.mark.gpu
def test_video_loader_matches_reference_loader_arithmetic(tmp_path):
    dev = "cuda:0"
    T, h0, w0 = 23, 120, 168                   # three chunks of 10 at chunk=10: exercises both staging slots
    rd, od = _write_video(str(tmp_path), "01", T, h0, w0, 3)
    _write_video(str(tmp_path), "02", 5, h0, w0, 4)
    ld = A.VideoLoader(dev, size=(64, 48), chunk=10, decode_threads=4)
    frames, flows = ld.load(A.loader.list_frames(rd), A.loader.list_frames(od))
    torch.cuda.synchronize()
    assert frames.shape == (T, 3, 48, 64) and flows.shape == (T - 1, 2, 48, 64)
    for t in (0, 9, 10, 22):
        ref = O.preprocess_frame(A.decode_frame(os.path.join(rd, "%04d.jpg" % t)), (64, 48))
        assert np.array_equal(frames[t].cpu().numpy(), ref), "frame %d" % t
    for t in (0, 10, 21):
        ref = O.preprocess_flow(A.read_flo(os.path.join(od, "%04d.flo" % t)), (64, 48))
        assert np.array_equal(flows[t].cpu().numpy(), ref), "flow %d" % t
    got = [(n, f.shape[0], o.shape[0]) for n, f, o in ld.iter_dataset(os.path.join(str(tmp_path), "rgb"), os.path.join(str(tmp_path), "op"))]
    assert got == [("01", T, T - 1), ("02", 5, 4)]
    # nvJPEG decode on the device: only compressed bytes cross PCIe; its IDCT differs from libjpeg-turbo by a grey level or two
    lg = A.VideoLoader(dev, size=(64, 48), chunk=10, gpu_jpeg=True)
    try:
        fg, _ = lg.load(A.loader.list_frames(rd), [])
    except Exception as e:                      # torchvision built without nvJPEG on this box: the mode is optional
        pytest.skip("nvJPEG decode unavailable: %s" % e)
    torch.cuda.synchronize()
    # nvJPEG upsamples chroma without libjpeg-turbo's "fancy" filter: large differences at chroma edges, small on average
    err = (fg - frames).abs()
    print("nvJPEG vs libjpeg-turbo decode: mean |diff| %.4f, max %.4f (range 2.0)" % (float(err.mean()), float(err.max())))
    assert float(err.mean()) <= 0.03
    ld.close(); lg.close()
    with pytest.raises(RuntimeError, match="CUDA"):
        A.VideoLoader("cpu")
