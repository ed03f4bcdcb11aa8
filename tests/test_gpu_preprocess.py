"""GPU parity: frame / flow preprocessing kernels against the oracle and the reference-generated fixture (bit-exact)."""
import hashlib

import numpy as np
import pytest
import torch

import ammc_oracle as O
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth
from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_preprocess_matches_reference_loaders_bit_exactly():
    c, g = load_golden("preprocess")
    for name, bgr, flow, size in synth.preprocess_inputs(c["seed"]):
        rgb = A.preprocess_frames(torch.from_numpy(bgr)[None].to(DEV), size)[0].cpu().numpy()
        op = A.preprocess_flow(torch.from_numpy(flow)[None].to(DEV), size)[0].cpu().numpy()
        meta = c["cases"][name]
        if "rgb_sha256" in meta:
            assert hashlib.sha256(rgb.tobytes()).hexdigest() == meta["rgb_sha256"], name
            assert hashlib.sha256(op.tobytes()).hexdigest() == meta["op_sha256"], name
        else:
            assert np.array_equal(rgb, g[name + "_rgb_out"]), name
            assert np.array_equal(op, g[name + "_op_out"]), name


@pytest.mark.parametrize("h0,w0,W,H,n", [(240, 360, 256, 256, 5), (45, 70, 256, 256, 3), (300, 200, 96, 160, 2),
                                         (256, 256, 256, 256, 70), (7, 3, 5, 11, 1)])
def test_preprocess_batches_against_oracle(h0, w0, W, H, n):
    """Batches (more frames than the grid's z extent included), up- and down-scaling: bit-exact against the oracle."""
    rng = np.random.default_rng(h0 * 1000 + w0)
    bgr = rng.integers(0, 256, (n, h0, w0, 3), dtype=np.uint8)
    flow = (rng.standard_normal((n, h0, w0, 2)) * 5).astype(np.float32)
    rgb = A.preprocess_frames(torch.from_numpy(bgr).to(DEV), (W, H)).cpu().numpy()
    op = A.preprocess_flow(torch.from_numpy(flow).to(DEV), (W, H)).cpu().numpy()
    for i in sorted({0, n // 2, n - 1}):
        assert np.array_equal(rgb[i], O.preprocess_frame(bgr[i], (W, H))), i
        assert np.array_equal(op[i], O.preprocess_flow(flow[i], (W, H))), i
    # properties that hold at any size: value range, identity resize is exact, channel 1 = channel 0 / W
    assert rgb.min() >= -1.0 and rgb.max() <= 1.0
    assert np.array_equal(op[:, 1], op[:, 0] / np.float32(W))
    if (h0, w0) == (H, W):
        ident = (bgr[..., ::-1].astype(np.float32) / np.float32(255) - np.float32(0.5)) / np.float32(0.5)
        assert np.array_equal(rgb, ident.transpose(0, 3, 1, 2))


def test_preprocess_feeds_the_scorer_and_refuses_bad_inputs():
    frames = [np.random.default_rng(i).integers(0, 256, (60, 90, 3), dtype=np.uint8) for i in range(9)]
    flows = [np.random.default_rng(100 + i).standard_normal((60, 90, 2)).astype(np.float32) for i in range(8)]
    rgb, op = A.load_video(frames, flows, DEV, (64, 64))
    assert rgb.shape == (9, 3, 64, 64) and op.shape == (8, 2, 64, 64)
    m = A.get_twostream()
    m.load_state_dict(synth.generator_params(4))
    rec = A.VideoScorer(m.to(DEV).eval(), batch=4).score_video(rgb, op)
    assert rec["rgb_img_pred"].shape == (9,) and np.isfinite(rec["rgb_img_pred"]).all()
    with pytest.raises(RuntimeError, match="CUDA uint8"):
        A.preprocess_frames(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="CUDA uint8"):
        A.preprocess_frames(torch.zeros(1, 8, 8, 3, device=DEV))
    with pytest.raises(RuntimeError, match="CUDA float32"):
        A.preprocess_flow(torch.zeros(1, 8, 8, 3, device=DEV))
    assert A.preprocess_frames(torch.zeros(0, 8, 8, 3, dtype=torch.uint8, device=DEV)).shape == (0, 3, 256, 256)
