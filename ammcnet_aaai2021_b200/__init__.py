"""B200-native (sm_100a) implementation of AMMC-Net's memory + AMFT + scoring hot path.

Public surface mirrors the reference (NjuHaoZhang/AMMCNet_AAAI2021, Code/models/unet.py, Code/utils/utils.py,
Code/main/eval_metric.py): `Quantize_topk`, `enc_quan_dec_topk`, `enc_quan_dec_res_topk`, `bridge`, `psnr_error`,
`evaluate`; plus `patch_reference` / `swap_modules` to slot them behind the reference's own factories.
Importing the package does not need a GPU; calling any op does (there is no CPU fallback).
"""
from .modules import Quantize_topk, enc_quan_dec_topk, enc_quan_dec_res_topk, bridge, double_conv, psnr_error
from .functions import psnr_per_frame
from .scoring import VideoScorer, assemble_video_records, score_reduce, evaluate, LAM_MAP
from .patch import patch_reference, unpatch_reference, swap_modules, patch_reference_losses, unpatch_reference_losses
from .host_model import twostream, UNetMem_v7, get_twostream, PixelDiscriminator, FlowNet2SD
from .graphs import GraphedPath
from .generator import GeneratorEngine
from .preprocess import preprocess_frames, preprocess_flow, load_video, widen_bf16, narrow_bf16
from .losses import (Intensity_Loss, Gradient_Loss, frame_losses, Flow_Loss, Adversarial_Loss, Discriminate_Loss,
                     Twostream_vq_Loss, Twostream_Loss, rgb_Loss, rgb_vq_Loss, op_loss, op_vq_Loss, op_loss_v1, op_vq_Loss_v1)
from .loader import VideoLoader, read_flo, write_flo, decode_frame

__all__ = [
    "Quantize_topk", "enc_quan_dec_topk", "enc_quan_dec_res_topk", "bridge", "double_conv", "psnr_error",
    "psnr_per_frame", "VideoScorer", "assemble_video_records", "score_reduce", "evaluate", "LAM_MAP",
    "patch_reference", "unpatch_reference", "swap_modules", "patch_reference_losses", "unpatch_reference_losses", "twostream", "UNetMem_v7", "get_twostream", "PixelDiscriminator", "FlowNet2SD", "GraphedPath", "GeneratorEngine", "preprocess_frames", "preprocess_flow", "load_video", "widen_bf16", "narrow_bf16", "Intensity_Loss", "Gradient_Loss", "frame_losses", "Flow_Loss", "Adversarial_Loss", "Discriminate_Loss",
    "Twostream_vq_Loss", "Twostream_Loss", "rgb_Loss", "rgb_vq_Loss", "op_loss", "op_vq_Loss", "op_loss_v1", "op_vq_Loss_v1", "VideoLoader", "read_flo", "write_flo", "decode_frame",
]
