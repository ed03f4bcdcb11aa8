"""CPU oracle for the AMMC-Net memory + AMFT + scoring hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this file, and only as the checker
(or, for `bench.py`, as the timed CPU baseline).  The shipped package `ammcnet_aaai2021_b200` never
imports it and has no CPU fallback.

It restates, op for op and in the same association order, what the reference computes with ATen on
the CPU (fp32; pass dtype=torch.float64 for an adjudicating high-precision run).  Every function cites
the reference lines it follows (paths relative to /root/reference).

Parity pinning: the reference's own tree holds no test for the memory module, AMFT or PSNR
(SURVEY.md section 4), so those parts are pinned by running the live reference in the build container
(`oracle/gen_golden.py` -> `tests/golden/*.npz`, re-checked by `tests/test_oracle_golden.py`).  The
score reduction is additionally pinned by the reference's recorded per-frame score pickles
(`Code/ammcnet_os/model_result_save/*`), whose records and reduced scores are committed as fixtures.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# memory module
# --------------------------------------------------------------------------------------------------
def squared_l2_distances(flatten: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """dist[n, j] = (||z_n||^2 - 2 z_n.e_j) + ||e_j||^2, same association as Code/models/unet.py:283-288."""
    return (flatten.pow(2).sum(1, keepdim=True) - 2 * flatten @ embed
            + embed.pow(2).sum(0, keepdim=True))


def quantize_topk_forward(z: torch.Tensor, embed: torch.Tensor, k: int, *, training: bool = False,
                          cluster_size: Optional[torch.Tensor] = None,
                          embed_avg: Optional[torch.Tensor] = None,
                          decay: float = 0.99, eps: float = 1e-5) -> Dict[str, torch.Tensor]:
    """Quantize_topk.forward, Code/models/unet.py:282-313 (embed_code: 315-316).

    z: [b, h, w, D]; embed: [D, M].  Returns the three outputs of the reference plus the indices, the
    distance matrix (for margin filters in tests) and -- in training -- the UPDATED buffers (the inputs
    are not mutated; outputs use the pre-update bank, unet.py:298-309 run after the gathers).
    """
    D, M = embed.shape
    flatten = z.reshape(-1, D)
    dist = squared_l2_distances(flatten, embed)
    embed_ind = (-dist).max(1)[1]                                     # unet.py:289
    bank_t = embed.transpose(0, 1)
    q1 = F.embedding(embed_ind.view(*z.shape[:-1]), bank_t)            # unet.py:291-292
    idx_topk = (-dist).topk(k, dim=1)[1]                               # unet.py:293 (sorted, nearest first)
    read = F.embedding(idx_topk.view(*z.shape[:-1], k), bank_t)        # [b,h,w,k,D]
    read = read.view(*z.shape[:-1], k * D)                             # unet.py:296
    out: Dict[str, torch.Tensor] = {}
    if training:
        onehot = F.one_hot(embed_ind, M).type(flatten.dtype)           # unet.py:290
        counts = onehot.sum(0)
        embed_sum = flatten.transpose(0, 1) @ onehot                   # unet.py:302
        new_cs = cluster_size * decay + (1 - decay) * counts           # unet.py:299-301
        new_avg = embed_avg * decay + (1 - decay) * embed_sum          # unet.py:303
        n = new_cs.sum()
        cs = (new_cs + eps) / (n + M * eps) * n                        # unet.py:304-307
        out.update(counts=counts, embed_sum=embed_sum, cluster_size=new_cs, embed_avg=new_avg,
                   embed=new_avg / cs.unsqueeze(0))                    # unet.py:308-309
    diff = (q1 - z).pow(2).mean()                                      # unet.py:310
    ste = z + (q1 - z)                                                 # unet.py:311 (value differs from q1 by <=1 ulp)
    out.update(quantize_topk=read, diff=diff, quantize=ste, quantize_raw=q1, idx_topk=idx_topk,
               idx_top1=embed_ind, dist=dist)
    return out


def memory_module_forward(x: torch.Tensor, enc_w: torch.Tensor, enc_b: torch.Tensor,
                          embed: torch.Tensor, dec_w: torch.Tensor, dec_b: torch.Tensor, k: int,
                          *, residual: bool = True, training: bool = False,
                          cluster_size=None, embed_avg=None, decay=0.99, eps=1e-5):
    """enc_quan_dec_topk.forward (unet.py:325-331) and, with residual=True, enc_quan_dec_res_topk.forward
    (unet.py:384-387).  x: [b, C, h, w] NCHW.  enc_w: [D, C, 1, 1] or [D, C]; dec_w: [C, kD, 1, 1] or [C, kD]."""
    D = enc_w.shape[0]
    C = dec_w.shape[0]
    z = F.conv2d(x, enc_w.reshape(D, -1, 1, 1), enc_b).permute(0, 2, 3, 1)   # unet.py:326
    q = quantize_topk_forward(z, embed, k, training=training, cluster_size=cluster_size,
                              embed_avg=embed_avg, decay=decay, eps=eps)
    read_nchw = q["quantize_topk"].permute(0, 3, 1, 2)                        # unet.py:328
    out = F.conv2d(read_nchw, dec_w.reshape(C, -1, 1, 1), dec_b)              # unet.py:330
    if residual:
        out = out + x                                                         # unet.py:386
    b = x.shape[0]
    sse_per_frame = (q["quantize_raw"] - z).pow(2).reshape(b, -1).sum(1)          # for per-frame partial parity
    res = dict(q)
    res.update(out=out, diff=q["diff"].unsqueeze(0), z=z, sse_per_frame=sse_per_frame)
    return res


def memory_module_backward(x, enc_w, enc_b, embed, dec_w, idx_topk, z, g_out, g_diff, g_q1=None,
                           residual: bool = True):
    """Closed-form backward of a9 (verified there against autograd of unet.py:282-331).

    The read is gathered from a buffer, so no gradient reaches z through it; the commit loss and the
    straight-through q1 do.  Returns dict(gx, g_enc_w, g_enc_b, g_dec_w, g_dec_b).
    """
    D, M = embed.shape
    C = dec_w.shape[0]
    k = idx_topk.shape[-1]
    b, _, h, w = x.shape
    N = b * h * w
    zf = z.reshape(N, D)
    bank_t = embed.t()
    e1 = bank_t[idx_topk.reshape(N, k)[:, 0]]
    gz = g_diff.reshape(()) * 2.0 * (zf - e1) / (N * D)
    if g_q1 is not None:
        gz = gz + g_q1.reshape(N, D)
    go = g_out.permute(0, 2, 3, 1).reshape(N, C)
    read = bank_t[idx_topk.reshape(N, k)].reshape(N, k * D)
    xf = x.permute(0, 2, 3, 1).reshape(N, -1)
    gx = gz @ enc_w.reshape(D, -1)
    if residual:
        gx = gx + go
    return dict(
        gx=gx.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous(),
        g_enc_w=(gz.t() @ xf).reshape(enc_w.shape), g_enc_b=gz.sum(0),
        g_dec_w=(go.t() @ read).reshape(dec_w.shape), g_dec_b=go.sum(0), gz=gz)


# --------------------------------------------------------------------------------------------------
# AMFT (class `bridge`)
# --------------------------------------------------------------------------------------------------
def double_conv_forward(u, p: Dict[str, torch.Tensor], prefix: str, training: bool = False,
                        momentum: float = 0.1, bn_eps: float = 1e-5):
    """double_conv.forward, Code/models/unet.py:8-20: (conv3x3 no-bias -> BN -> ReLU) x 2.
    `p` is a state_dict; keys `<prefix>.conv.{0,3}.weight`, `<prefix>.conv.{1,4}.{weight,bias,running_mean,running_var}`.
    In training mode batch statistics are used and the updated running stats are returned."""
    new_stats = {}
    for ci, bi in ((0, 1), (3, 4)):
        u = F.conv2d(u, p[f"{prefix}.conv.{ci}.weight"], None, padding=1)
        rm, rv = p[f"{prefix}.conv.{bi}.running_mean"], p[f"{prefix}.conv.{bi}.running_var"]
        if training:
            rm, rv = rm.clone(), rv.clone()
        u = F.batch_norm(u, rm, rv, p[f"{prefix}.conv.{bi}.weight"], p[f"{prefix}.conv.{bi}.bias"],
                         training=training, momentum=momentum, eps=bn_eps)
        if training:
            new_stats[f"{prefix}.conv.{bi}.running_mean"] = rm
            new_stats[f"{prefix}.conv.{bi}.running_var"] = rv
        u = F.relu(u)
    return u, new_stats


def amft_forward(zx, zy, p: Dict[str, torch.Tensor], prefix: str = "", training: bool = False):
    """bridge.forward, Code/models/unet.py:962-965: x' = zx + O2F(zy), y' = zy + F20(zx)."""
    pre = prefix + "." if prefix and not prefix.endswith(".") else prefix
    fo, s1 = double_conv_forward(zy, p, pre + "O2F", training)
    ff, s2 = double_conv_forward(zx, p, pre + "F20", training)
    s1.update(s2)
    return zx + fo, zy + ff, s1


# --------------------------------------------------------------------------------------------------
# PSNR + per-frame records + score reduction
# --------------------------------------------------------------------------------------------------
def psnr_per_frame(gen: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """Per-sample value of psnr_error (Code/utils/utils.py:141-147) -- the loop at
    Code/run_helper/test_helper.py:445-452 calls it with batch 1, so the mean at utils.py:148 is the identity."""
    n = gen.shape[1] * gen.shape[2] * gen.shape[3]
    sq = ((gt + 1.0) / 2.0 - (gen + 1.0) / 2.0) ** 2
    return 10 * torch.log10(1.0 / ((1.0 / n) * torch.sum(sq, [1, 2, 3])))


def psnr_error(gen: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """psnr_error, Code/utils/utils.py:130-148 (mean over the batch axis)."""
    return torch.mean(psnr_per_frame(gen, gt))


def assemble_video_records(psnr_clip: np.ndarray, commit_group: np.ndarray, clip_len: int = 5,
                           group: int = 16, tail_copy: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """Record semantics of Code/run_helper/test_helper.py:414-475 for ONE sub-video.

    psnr_clip[c]  : PSNR of clip c (predicting frame c + clip_len - 1), c = 0..n_clips-1
    commit_group[g]: the batch-level commit scalar of the g-th group of `group` consecutive clips
    Returns (img_pred_arr, fea_comm_arr) of length n_clips + clip_len - 1; the first clip_len-1 entries
    copy entry clip_len-1 (test_helper.py:465,472); with tail_copy (the op stream, whose clip length is
    one shorter than the frame count implies) the last entry copies the one before (test_helper.py:467,474).
    """
    n_clips = len(psnr_clip)
    num_frame = n_clips + clip_len - 1 + (1 if tail_copy else 0)
    img = np.empty((num_frame,), np.float32)
    fea = np.empty((num_frame,), np.float32)
    for c in range(n_clips):
        img[c + clip_len - 1] = psnr_clip[c]
        fea[c + clip_len - 1] = commit_group[c // group]
    img[:clip_len - 1] = img[clip_len - 1]
    fea[:clip_len - 1] = fea[clip_len - 1]
    if tail_copy:
        img[num_frame - 1] = img[num_frame - 2]
        fea[num_frame - 1] = fea[num_frame - 2]
    return img, fea


DECIDABLE_IDX = 4  # Code/main/eval_metric.py:15-17


def norm_score(records: Sequence[np.ndarray]) -> np.ndarray:
    """norm_score, Code/main/eval_metric.py:405-417 (float32; does not mutate its input, unlike the
    reference which normalises the record arrays in place)."""
    scores = np.array([], dtype=np.float32)
    for rec in records:
        d = np.array(rec, dtype=np.float32, copy=True)
        d -= d.min()
        d /= d.max()
        scores = np.concatenate((scores, d[DECIDABLE_IDX:]), axis=0)
    scores -= scores.min()
    scores /= scores.max()
    return scores


def score_reduce(img_records: Sequence[np.ndarray], fea_records: Sequence[np.ndarray],
                 lam: Tuple[float, float]) -> np.ndarray:
    """Regularity scores, Code/main/eval_metric.py:418-427: mix then the non-recursive 2-tap smoothing
    that runs across video boundaries.  Returns float32 [T]."""
    img = norm_score(img_records)
    fea = norm_score(fea_records)
    identity = np.ones_like(fea)
    l1, l2 = lam[0], lam[1]
    s = (1 - l1) * img + l1 * (identity - fea)
    out = [(1 - l2) * s[i - 1] + l2 * s[i] if i > 0 else s[i] for i in range(len(s))]
    return np.asarray(out, dtype=np.float32)


def roc_auc(labels: np.ndarray, scores: np.ndarray, pos_label: int = 0) -> float:
    """sklearn.metrics.roc_curve + auc as called at Code/main/eval_metric.py:428-429, restated with numpy
    (trapezoid over the ROC polyline with tied scores merged) so the oracle does not need sklearn."""
    y = (np.asarray(labels) == pos_label).astype(np.float64)
    s = np.asarray(scores, dtype=np.float64)
    order = np.argsort(-s, kind="mergesort")
    y, s = y[order], s[order]
    distinct = np.where(np.diff(s))[0]
    thr = np.r_[distinct, y.size - 1]
    tps = np.cumsum(y)[thr]
    fps = 1 + thr - tps
    tps = np.r_[0, tps]
    fps = np.r_[0, fps]
    if tps[-1] == 0 or fps[-1] == 0:
        return float("nan")
    return float(np.trapezoid(tps / tps[-1], fps / fps[-1]))


# --------------------------------------------------------------------------------------------------
# whole path (one "step" of the metric): both memory modules + AMFT + rgb PSNR
# --------------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------------
# the U-Net encoder / decoder around the path, eval mode
# --------------------------------------------------------------------------------------------------
def unet_encode(x, p: Dict[str, torch.Tensor], s: str):
    """inc / down1-3 of UNetMem_v7 (Code/models/unet.py:924-928; inconv 23-30, down 33-42): x1, x2, x3, x4."""
    x1, _ = double_conv_forward(x, p, s + "inc.conv")
    x2, _ = double_conv_forward(F.max_pool2d(x1, 2), p, s + "down1.mpconv.1")
    x3, _ = double_conv_forward(F.max_pool2d(x2, 2), p, s + "down2.mpconv.1")
    x4, _ = double_conv_forward(F.max_pool2d(x3, 2), p, s + "down3.mpconv.1")
    return x1, x2, x3, x4


def unet_up(x1, x2, p: Dict[str, torch.Tensor], prefix: str):
    """up.forward, Code/models/unet.py:50-59: ConvTranspose2d(2, stride 2), pad to the skip, cat([skip, up]), double_conv."""
    x1 = F.conv_transpose2d(x1, p[prefix + ".up.weight"], p[prefix + ".up.bias"], stride=2)
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
    out, _ = double_conv_forward(torch.cat([x2, x1], dim=1), p, prefix + ".conv")
    return out


def unet_decode(x4, x3, x2, x1, p: Dict[str, torch.Tensor], s: str):
    """up1-3 + outc of UNetMem_v7 (Code/models/unet.py:930-933); the caller applies tanh (unet.py:937)."""
    x = unet_up(x4, x3, p, s + "up1")
    x = unet_up(x, x2, p, s + "up2")
    x = unet_up(x, x1, p, s + "up3")
    return F.conv2d(x, p[s + "outc.weight"], p[s + "outc.bias"], padding=1)


def twostream_forward(rgb_x, op_x, p: Dict[str, torch.Tensor], k: int):
    """twostream.forward in eval mode, Code/models/unet.py:981-1007.
    Returns (tanh rgb, tanh op, (rgb_diff, op_diff), (rgb_q1, op_q1))."""
    enc, mem = {}, {}
    for s, x in (("rgb", rgb_x), ("op", op_x)):
        enc[s] = unet_encode(x, p, s + ".")
        pre = f"{s}.vq_down3.quan."
        mem[s] = memory_module_forward(enc[s][3], p[pre + "enc.weight"], p[pre + "enc.bias"], p[pre + "quantize.embed"],
                                       p[pre + "dec.weight"], p[pre + "dec.bias"], k)
    r4, o4, _ = amft_forward(mem["rgb"]["out"], mem["op"]["out"], p, "bridge")
    rgb_y = unet_decode(r4, enc["rgb"][2], enc["rgb"][1], enc["rgb"][0], p, "rgb.")
    op_y = unet_decode(o4, enc["op"][2], enc["op"][1], enc["op"][0], p, "op.")
    return (torch.tanh(rgb_y), torch.tanh(op_y), (mem["rgb"]["diff"], mem["op"]["diff"]),
            (mem["rgb"]["quantize"], mem["op"]["quantize"]))


# --------------------------------------------------------------------------------------------------
# frame / flow preprocessing in front of the generator
# --------------------------------------------------------------------------------------------------
# The reference resizes with cv2.resize (default INTER_LINEAR; opencv-python==4.1.1.26, Code/environment.yaml:78), a
# third-party dependency that is not under /root/reference.  Its published algorithm (modules/imgproc/src/resize.cpp,
# resizeGeneric_ / HResizeLinear / VResizeLinear) is restated here and pinned bit-exactly against the cv2 of this image
# and against the reference's own _load_frame / _load_op by oracle/gen_golden.py.
def _linear_taps(src: int, dst: int, clamp_weights: bool):
    """Source indices and float32 weights of cv2's bilinear resize along one axis.
    scale = 1 / (dst / src) in double; f = float((d + 0.5) * scale - 0.5); s = floor(f); f -= s.  Along x the weights are
    clamped at the borders (f = 0), along y only the row indices are (resize.cpp: xofs/alpha vs. yofs/beta + clip)."""
    scale = 1.0 / (float(dst) / float(src))
    i0 = np.empty(dst, np.int64)
    i1 = np.empty(dst, np.int64)
    wgt = np.empty((dst, 2), np.float32)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp_weights:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= src - 1:
                f, s = np.float32(0), src - 1
        wgt[d, 0] = np.float32(1.0) - f
        wgt[d, 1] = f
        i0[d] = min(max(s, 0), src - 1)
        i1[d] = min(max(s + 1, 0), src - 1)
    return i0, i1, wgt


def resize_linear_u8(img: np.ndarray, width: int, height: int) -> np.ndarray:
    """cv2.resize(img, (width, height)) for uint8 [h, w, c]: 11-bit fixed-point weights (cvRound(w * 2048)), integer
    horizontal pass, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2."""
    h0, w0 = img.shape[:2]
    x0, x1, ax = _linear_taps(w0, width, True)
    y0, y1, ay = _linear_taps(h0, height, False)
    iax = np.rint(ax * np.float32(2048)).astype(np.int64)
    iay = np.rint(ay * np.float32(2048)).astype(np.int64)
    src = img.astype(np.int64).reshape(h0, w0, -1)
    rows = src[:, x0, :] * iax[:, 0][None, :, None] + src[:, x1, :] * iax[:, 1][None, :, None]
    s0, s1 = rows[y0], rows[y1]
    b0, b1 = iay[:, 0][:, None, None], iay[:, 1][:, None, None]
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((height, width) + img.shape[2:])


def resize_linear_f32(img: np.ndarray, width: int, height: int) -> np.ndarray:
    """cv2.resize for float32 [h, w, c]: float32 weights, separate multiply and add (no FMA), horizontal then vertical."""
    h0, w0 = img.shape[:2]
    x0, x1, ax = _linear_taps(w0, width, True)
    y0, y1, ay = _linear_taps(h0, height, False)
    src = img.astype(np.float32).reshape(h0, w0, -1)
    rows = (src[:, x0, :] * ax[:, 0][None, :, None]).astype(np.float32) + (src[:, x1, :] * ax[:, 1][None, :, None]).astype(np.float32)
    s0, s1 = rows[y0], rows[y1]
    out = (s0 * ay[:, 0][:, None, None]).astype(np.float32) + (s1 * ay[:, 1][:, None, None]).astype(np.float32)
    return out.astype(np.float32).reshape((height, width) + img.shape[2:])


def preprocess_frame(bgr_u8: np.ndarray, size=(256, 256)) -> np.ndarray:
    """_load_frame with the test transform (Code/dataset/two_stream_dataset.py:72-84, 501-505): BGR->RGB, resize,
    ToTensor (uint8 HWC -> float32 CHW / 255), Normalize(0.5, 0.5).  Returns float32 [3, H, W]."""
    width, height = size
    rgb = resize_linear_u8(bgr_u8[:, :, ::-1], width, height)
    x = rgb.astype(np.float32) / np.float32(255)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def preprocess_flow(flow: np.ndarray, size=(256, 256)) -> np.ndarray:
    """_load_op (Code/dataset/two_stream_dataset.py:86-99): resize, ch0 = ch0 * 1.0 / H, ch1 = (scaled ch0) / W --
    the loader's quirk: channel 1 is derived from channel 0, the resized v component is discarded.  float32 [2, H, W]."""
    width, height = size
    r = resize_linear_f32(flow.astype(np.float32), width, height)
    c0 = (r[:, :, 0] * np.float32(1.0)) / np.float32(height)
    c1 = c0 / np.float32(width)
    return np.stack([c0, c1]).astype(np.float32)


# --------------------------------------------------------------------------------------------------
# image-space generator losses (the part that needs no external weights)
# --------------------------------------------------------------------------------------------------
def intensity_loss(gen: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """Intensity_Loss(l_num=2) -> L2, Code/models/losses/losses_utils.py:17-28,124-129: the L2 norm over the channel axis,
    averaged over batch and pixels."""
    return torch.norm(gen - gt, p=2, dim=1).mean()


def gradient_loss(gen: torch.Tensor, gt: torch.Tensor, alpha: int = 1) -> torch.Tensor:
    """Gradient_Loss, losses_utils.py:30-59: the [-1, 1] difference filters are [1, C, 1, 2] / [1, C, 2, 1] conv weights, i.e.
    they SUM over the channels (one output channel); zero padding on the left / top; mean of |d_x|^a + |d_y|^a."""
    C = gen.shape[1]
    filt = torch.tensor([[-1.0, 1.0]], dtype=gen.dtype, device=gen.device)
    fx = filt.view(1, 1, 1, 2).repeat(1, C, 1, 1)
    fy = filt.view(1, 1, 2, 1).repeat(1, C, 1, 1)
    gen_dx = F.conv2d(F.pad(gen, (1, 0, 0, 0)), fx)
    gen_dy = F.conv2d(F.pad(gen, (0, 0, 1, 0)), fy)
    gt_dx = F.conv2d(F.pad(gt, (1, 0, 0, 0)), fx)
    gt_dy = F.conv2d(F.pad(gt, (0, 0, 1, 0)), fy)
    return torch.mean(torch.abs(gt_dx - gen_dx) ** alpha + torch.abs(gt_dy - gen_dy) ** alpha)


def flow_loss(gen_flows: torch.Tensor, gt_flows: torch.Tensor) -> torch.Tensor:
    """Flow_Loss, Code/models/losses/losses_utils.py:10-15."""
    return torch.mean(torch.abs(gen_flows - gt_flows))


def adversarial_loss(fake_outputs: torch.Tensor) -> torch.Tensor:
    """Adversarial_Loss, losses_utils.py:103-107 (least-squares GAN, generator side)."""
    return torch.mean((fake_outputs - 1) ** 2 / 2)


def discriminate_loss(real_outputs: torch.Tensor, fake_outputs: torch.Tensor) -> torch.Tensor:
    """Discriminate_Loss, losses_utils.py:109-113 (least-squares GAN, discriminator side)."""
    return torch.mean((real_outputs - 1) ** 2 / 2) + torch.mean(fake_outputs ** 2 / 2)


def twostream_vq_loss(lam: Dict[str, float], flow_pred, flow_gt, rgb_out, rgb_target, op_out, op_target, latent_diff, d_gen):
    """Twostream_vq_Loss.forward, Code/models/losses/loss_zoo.py:312-350: the weighted generator objective of the joint
    training step.  `lam` holds lam_adv, lam_gdl, lam_flow, lam_lp, lam_latent, lam_lp_op (base_Loss, loss_zoo.py:15-31).
    Returns (g_loss, dict of the seven scalars the reference stores as attributes)."""
    g_adv = adversarial_loss(d_gen)
    g_flow = flow_loss(flow_pred, flow_gt)
    g_int = intensity_loss(rgb_out, rgb_target)
    g_gd = gradient_loss(rgb_out, rgb_target)
    g_int_op = intensity_loss(op_out, op_target)
    g_loss = lam["lam_adv"] * g_adv + lam["lam_gdl"] * g_gd + lam["lam_flow"] * g_flow + lam["lam_lp"] * g_int + \
        lam["lam_latent"] * latent_diff + lam["lam_lp_op"] * g_int_op
    parts = dict(g_loss=g_loss, g_adv_loss=g_adv, g_flow_loss=g_flow, g_int_loss=g_int, g_gd_loss=g_gd, g_int_loss_op=g_int_op,
                 g_latent_loss=latent_diff)
    return g_loss, parts


def path_forward(x_rgb, x_op, gen, gt, params: Dict[str, torch.Tensor], k: int):
    """The starred region of twostream.forward (Code/models/unet.py:985-994) followed by the per-frame rgb
    PSNR of test_helper.py:445-452.  `params` uses the reference state_dict key names
    (rgb.vq_down3.quan.enc.weight, ..., bridge.O2F.conv.0.weight, ...)."""
    outs = {}
    for s, x in (("rgb", x_rgb), ("op", x_op)):
        pre = f"{s}.vq_down3.quan."
        outs[s] = memory_module_forward(x, params[pre + "enc.weight"], params[pre + "enc.bias"],
                                        params[pre + "quantize.embed"], params[pre + "dec.weight"],
                                        params[pre + "dec.bias"], k)
    yx, yy, _ = amft_forward(outs["rgb"]["out"], outs["op"]["out"], params, "bridge")
    return dict(rgb=outs["rgb"], op=outs["op"], amft_rgb=yx, amft_op=yy,
                psnr=psnr_per_frame(gen, gt))
