/*
 * ammc_b200.h -- C ABI of the B200-native AMMC-Net memory + AMFT + scoring hot path.
 *
 * The reference (NjuHaoZhang/AMMCNet_AAAI2021) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md section 2.2): its "plugin API" for this path is a set of torch.nn.Module classes and two
 * functions.  Each entry point below replaces the ATen op sequence of one of them; the Python mirror in
 * ammcnet_aaai2021_b200/ binds these symbols with ctypes and keeps the reference's class names,
 * signatures and state_dict keys.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; tensors are dense, row-major in the
 *     order written in the comment; fp32 unless stated; indices are int64 like torch's.
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing synchronises,
 *     nothing allocates: the caller owns inputs, outputs and the workspace (size from the *_workspace_bytes
 *     query) and may free them once the stream has passed the call.
 *   - return value: 0 on success, negative AMMC_E* on failure; ammc_last_error() gives the message of the
 *     calling thread's last failure.  There is no CPU fallback and no alternate backend.
 *   - notation: N = b*h*w pixels ("queries"), C feature channels, D embed_dim, M n_embed (memory items), k.
 */
#ifndef AMMC_B200_H_
#define AMMC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMMC_OK 0
#define AMMC_EINVAL (-1)      /* bad pointer / size / unsupported shape */
#define AMMC_ECUDA (-2)       /* CUDA runtime or driver error (message has the cudaError string) */
#define AMMC_EWORKSPACE (-3)  /* workspace too small */
#define AMMC_EUNSUPPORTED (-4)/* shape outside what the sm_100a kernels implement */

#define AMMC_MAX_K 8          /* top-k width supported by the addressing kernels */

int ammc_version(void);
const char* ammc_last_error(void);
/* 1 when the current device is sm_100 (B200); the library refuses to run anywhere else. */
int ammc_device_supported(void);
/* Every mbarrier wait in the tcgen05 kernels is bounded (~10 s of spinning, far beyond any legitimate wait): a pipeline that
 * stalls records where and executes `trap`, so the launch fails, the CUDA context carries a sticky error and every later
 * ammc_* call returns AMMC_ECUDA -- a broken pipeline can neither hang the GPU nor let later launches run on garbage.
 * ammc_pipeline_check synchronises the device and returns 0 when healthy, AMMC_ECUDA after a trap (1 with out4 = {kernel
 * family, wait tag, block, thread} if the record could still be read).  Hardware-behaviour probes live in
 * include/ammc_b200_debug.h and are compiled only into the separate debug library. */
int ammc_pipeline_check(int* out4);

/* ---------------------------------------------------------------------------------------------------
 * Memory module.  Replaces enc_quan_dec_topk.forward / enc_quan_dec_res_topk.forward
 * (reference Code/models/unet.py:325-331, 384-387) including Quantize_topk.forward (unet.py:282-313).
 *
 *   x        [b, C, h, w]   NCHW input features
 *   enc_w    [D, C], enc_b [D]          1x1 conv `enc`  (unet.py:321)
 *   embed    [D, M]                     memory bank buffer (unet.py:277-278), column j = item j
 *   dec_w    [C, k*D], dec_b [C]        1x1 conv `dec`  (unet.py:323)
 *   out      [b, C, h, w]   dec(read) + dec_b (+ x when residual != 0)
 *   q1       [N, D]         value of the straight-through top-1 read  z + (e_top1 - z)   (unet.py:311)
 *   idx      [N, k] int64   top-k item indices, nearest first                                (unet.py:293)
 *   z        [N, D]         enc output in NHWC order (saved for backward / EMA)              (unet.py:326)
 *   sse_frame[b]            per-frame sum of (e_top1 - z)^2   (the per-frame partial of unet.py:310)
 *   diff     [1]            mean over all N*D elements = reference `diff`                    (unet.py:310,329)
 *   counts   [M], embed_sum [D, M]  (both NULL in eval) assignment statistics of unet.py:298-302
 *   out_planes NULL, or `out` additionally as the AMFT block's NHWC operand, written by the same epilogue:
 *              planes_fmt 0: [2][b,h,w,C] bf16 hi/lo planes; planes_fmt 1: a q buffer (ammc_q_act_bytes(N*C); its scale is
 *              derived on the device from max|x| -- reduced inside the enc kernel -- and a Cauchy-Schwarz bound on dec(read)).
 *              Only allowed when ammc_mem_dec_uses_tensor(...) == 1.
 * `dec` runs as a split-bf16 x3 GEMM on tcgen05 when k*D % 64 == 0, C % 64 == 0 and the feature map is at most 128
 * pixels wide; otherwise as an exact fp32 gather of precomputed table rows.  ammc_set_dec_mode: 0 auto,
 * 1 force the fp32 gather, 2 force the tensor-core GEMM.
 * ------------------------------------------------------------------------------------------------- */
int ammc_mem_dec_uses_tensor(int b, int h, int w, int C, int D, int M, int k);
/* Everything ammc_mem_fwd derives from the parameters alone -- packed bf16 planes of enc.weight / dec.weight, the bank as
 * rows with its norms and bf16 copy, the bound behind the q-plane scale -- can be prepared once per parameter version:
 * ammc_mem_prepare fills a caller-owned buffer of ammc_mem_prep_bytes(C, D, M, k) bytes (b, h, w select the same kernel
 * paths as the forward that will use it); passing it as `prep` to ammc_mem_fwd skips those launches.  prep == NULL keeps
 * the forward self-contained (the same quantities are derived into the workspace).  The reference re-derives them on
 * every call too (embed.t(), unet.py:316; the conv weights are cuDNN's business there).
 * With embed_dim == 64, n_embed <= 256, h*w % 128 == 0 and the tensor-core enc / dec paths in use, the eval forward runs
 * enc + addressing + exact refine as ONE persistent kernel (csrc/mem_front.cu); ammc_set_front_mode(0) restores the staged
 * kernels (A/B measurements). */
size_t ammc_mem_prep_bytes(int C, int D, int M, int k);
int ammc_mem_prepare(const float* enc_w, const float* embed, const float* dec_w, const float* dec_b, void* prep,
                     size_t prep_bytes, int b, int h, int w, int C, int D, int M, int k, void* stream);
int ammc_set_front_mode(int on);
int ammc_set_dec_mode(int mode);
/* `enc` runs as a split-bf16 x3 GEMM on tcgen05 that converts the fp32 NCHW input on the fly (and emits bf16(z) and
 * ||z||^2 for the addressing filter) when embed_dim == 64, C % 64 == 0 and h*w % 128 == 0; otherwise as an fp32 FFMA
 * GEMM.  ammc_set_enc_mode: 0 auto, 1 force fp32 FFMA, 2 force the tensor-core kernel. */
int ammc_set_enc_mode(int mode);
size_t ammc_mem_workspace_bytes(int b, int h, int w, int C, int D, int M, int k);

/* Addressing path (process-wide): 0 = auto (tensor-core filter + exact fp32 refine when D % 64 == 0, 16 <= M <= 65536, k <= 4,
 * else the generic fp32 CUDA-core kernel), 1 = force the generic fp32 kernel, 2 = force the tensor-core path (calls
 * with unsupported shapes then fail).  Both paths return bit-identical indices: the fp16 tensor-core pass (operands scaled
 * by powers of two, per query row / per bank) only pre-selects candidates (every item within a rigorous error margin of the
 * k-th best approximate score), which are
 * re-ranked with the exact fp32 arithmetic of the generic kernel; queries whose candidate list overflowed are
 * re-scanned exactly.  After ammc_mem_fwd / ammc_quantize_fwd the
 * first 8 bytes of the workspace hold two int32: [0] queries that needed the exact re-scan, [1] path used (1 or 2). */
int ammc_set_addressing_mode(int mode);

int ammc_mem_fwd(const float* x, const float* enc_w, const float* enc_b, const float* embed,
                 const float* dec_w, const float* dec_b,
                 float* out, float* q1, int64_t* idx, float* z, float* sse_frame, float* diff,
                 float* counts, float* embed_sum, void* out_planes, int planes_fmt, const void* prep,
                 void* workspace, size_t workspace_bytes,
                 int b, int h, int w, int C, int D, int M, int k, int residual, void* stream);

/* bf16 feature-I/O variant of the module (BASELINE configs[2]; the reference module is dtype-agnostic, unet.py:318-331,
 * 379-387): x and out are bf16 NCHW, everything else as ammc_mem_fwd.  x enters the enc GEMM exactly (bf16 is its own hi
 * plane), the residual sum is formed in fp32 and rounded once when `out` is stored, so indices, q1, z, diff and the
 * operand planes are bit-identical to ammc_mem_fwd on the widened tensor and out == bf16(its out).  Runs on the fused
 * front kernel + tensor-core dec only (ammc_mem_io16_supported); eval (counts == embed_sum == NULL). */
int ammc_mem_io16_supported(int b, int h, int w, int C, int D, int M, int k);
int ammc_mem_fwd_io16(const void* x_bf16, const float* enc_w, const float* enc_b, const float* embed,
                      const float* dec_w, const float* dec_b,
                      void* out_bf16, float* q1, int64_t* idx, float* z, float* sse_frame, float* diff,
                      void* out_planes, int planes_fmt, const void* prep,
                      void* workspace, size_t workspace_bytes,
                      int b, int h, int w, int C, int D, int M, int k, int residual, void* stream);
/* fp32 -> bf16 (round to nearest even), the inverse direction of ammc_cast_bf16_f32 */
int ammc_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

/* Staged form of the tensor-core addressing filter (composed internally by ammc_mem_fwd / ammc_quantize_fwd; exposed
 * for the addressing microbench of BASELINE configs[4]):
 *   ammc_addr_padded_items(M)   items after padding to the MMA N tile (Mpad)
 *   ammc_addr_pack_queries      z [N,D] fp32 -> zp [N,D] fp16 of the row scaled by a power of two s_n (max component in
 *                               [2^14, 2^15): no overflow, no lost small rows), zmeta [N,2] = (||z_n||^2, 1 / s_n)
 *   ammc_addr_pack_bank         embed [D,M] -> bank_t [M,D] fp32, en2 [M], bank_hi [Mpad,D] fp16 of the bank scaled by one power
 *                               of two t, en2pad [Mpad], emax [4] = {max ||e||, 1 / t, scratch, -}
 *   ammc_addr_filter            zp x bank_hi on tcgen05 (kind::f16, fp32 accumulate) -> per query the merged candidate list
 *                               cand [N,24] int32 and cand_cnt [N,2]:
 *                                 cand_cnt[n][0] = entries used.  The list is a guaranteed superset of the exact top-k:
 *                                 every item whose approximate score lies within the margin (8 * 2^-11 * ||z|| max||e|| +
 *                                 an fp32 evaluation slack) of the k-th smallest.  > 24 = a list overflowed -> the tail
 *                                 re-scans that query exactly over all items;
 *                                 cand_cnt[n][1] = 1 when the filter DECIDED the row: the k best approximate scores (and
 *                                 the next survivor) lie more than the margin apart, so cand[n][0..k) are the final
 *                                 indices in rank order and no exact distance is needed; 0 = the tail ranks the
 *                                 candidates by their exact fp32 distances */
#define AMMC_ADDR_CAND 24
int ammc_addr_padded_items(int M);
int ammc_addr_pack_queries(const float* z, void* zp, float* zmeta, int64_t N, int D, void* stream);
int ammc_addr_pack_bank(const float* embed, float* bank_t, float* en2, void* bank_hi, float* en2pad, float* emax,
                        int D, int M, void* stream);
int ammc_addr_filter(const void* zp, const float* zmeta, const void* bank_hi, const float* en2pad, const float* emax,
                     int* cand, int* cand_cnt, int64_t N, int D, int M, int k, void* stream);

/* Quantize_topk.forward on its own (unet.py:282-313): z [N, D] contiguous (N = frames*rows_per_frame).
 *   read [N, k*D]  concatenated top-k items, nearest first (unet.py:295-297); other outputs as above. */
size_t ammc_quantize_workspace_bytes(int64_t N, int D, int M, int k);

int ammc_quantize_fwd(const float* z, const float* embed,
                      float* read, float* q1, int64_t* idx, float* sse_frame, float* diff,
                      float* counts, float* embed_sum,
                      void* workspace, size_t workspace_bytes,
                      int64_t N, int64_t rows_per_frame, int D, int M, int k, void* stream);

/* Backward of Quantize_topk.forward w.r.t. its input (autograd of unet.py:310-311):
 *   gz[n,:] = g_diff * 2 (z_n - e_top1(n)) / (N*D) + g_q1[n,:]        (g_q1 may be NULL) */
size_t ammc_quantize_bwd_workspace_bytes(int64_t N, int D, int M);

int ammc_quantize_bwd(const float* z, const float* embed, const int64_t* idx, const float* g_diff,
                      const float* g_q1, float* gz, void* workspace, size_t workspace_bytes,
                      int64_t N, int D, int M, int k, void* stream);

/* Quantize_topk.embed_code (unet.py:315-316): out[i, :] = embed[:, ids[i]]. */
int ammc_embed_code(const int64_t* ids, const float* embed, float* out, int64_t n_ids, int D, int M,
                    void* stream);

/* Training-mode bank update (unet.py:298-309), in place on the registered buffers:
 *   cluster_size <- decay*cluster_size + (1-decay)*counts ;  embed_avg <- decay*embed_avg + (1-decay)*embed_sum
 *   embed <- embed_avg / ((cluster_size+eps)/(sum+M*eps)*sum)
 * In data-parallel training all-reduce counts / embed_sum over ranks BEFORE this call. */
int ammc_ema_update(float* embed, float* cluster_size, float* embed_avg,
                    const float* counts, const float* embed_sum,
                    int D, int M, float decay, float eps, void* stream);

/* Backward of the memory module (autograd of unet.py:282-331, 384-387).
 *   g_out [b,C,h,w], g_diff [1] (device), g_q1 [N,D] or NULL
 *   gx [b,C,h,w], g_enc_w [D,C], g_enc_b [D], g_dec_w [C,k*D], g_dec_b [C]   (all overwritten)
 * The read is gathered from a buffer, so no gradient reaches z through it and none reaches embed. */
size_t ammc_mem_bwd_workspace_bytes(int b, int h, int w, int C, int D, int M, int k);

int ammc_mem_bwd(const float* x, const float* enc_w, const float* embed, const int64_t* idx,
                 const float* z, const float* g_out, const float* g_diff, const float* g_q1,
                 float* gx, float* g_enc_w, float* g_enc_b, float* g_dec_w, float* g_dec_b,
                 void* workspace, size_t workspace_bytes,
                 int b, int h, int w, int C, int D, int M, int k, int residual, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * AMFT.  Replaces bridge.forward (unet.py:962-965) = two double_conv stacks (unet.py:8-20) + residuals.
 * The block is run as four 3x3 implicit-GEMM convolutions on tcgen05 tensor cores over NHWC bf16
 * operand planes.  `precision`: 1 = single bf16 pass; 3 = three-pass split-bf16 (hi*hi + hi*lo + lo*hi,
 * fp32 accumulate, ~2^-17 relative error: the fp32-parity mode); 2 = fp16 main product + e4m3 cross terms on q buffers
 * (below): the same parity bar at two pass-equivalents -- xp / wp / out_planes are then q buffers.
 *
 * ammc_pack_conv_weights:   w [Cout, Cin, 3, 3] fp32  ->  wp [2 planes][Cout][9*Cin] bf16 (k = tap*Cin+cin)
 * ammc_pack_nhwc:           x [b, C, h, w] fp32 NCHW  ->  xp [2 planes][b, h, w, C] bf16   (hi, lo)
 * ammc_conv3x3_bn_relu:     y = relu(conv3x3(x) * scale[c] + shift[c])  (+ res)   BN folded to scale/shift:
 *        scale = gamma / sqrt(var + eps), shift = beta - mean * scale       (unet.py:11-16, eval mode)
 *     out_planes != NULL : y written as NHWC bf16 (hi, lo) planes (input of the next conv)
 *     out_nchw   != NULL : y (+ res_nchw if not NULL) written as fp32 NCHW
 * ------------------------------------------------------------------------------------------------- */
int ammc_pack_conv_weights(const float* w, void* wp, int Cout, int Cin, void* stream);
int ammc_pack_nhwc(const float* x, void* xp, int b, int C, int h, int w, void* stream);
/* ---- precision 2: the "q" operand format (fp32 parity at two tensor-core pass-equivalents instead of three) ----------
 * A tensor t with a power-of-two scale s16 (|t*s16| < 2^15, derived on the device from max|t| or from a rigorous bound)
 * is stored as  h16 = fp16(t*s16),  h8 = e4m3(t*s16 / 128),  l8 = e4m3((t*s16 - h16) * 16).  The conv accumulates the
 * cross terms h8.l8 + l8.h8 with kind::f8f6f4 MMAs (K = 32: twice the MACs per cycle) and then h16.h16 with kind::f16
 * MMAs whose first instruction rescales the cross-term sum by 2^-4 (scale-input-d) -- one TMEM accumulator, each
 * operand byte read once, ~5e-5 relative error per conv (tests/test_precision_schemes.py models it on the CPU).
 *   activation buffer (n = b*h*w*C, NHWC):  [n fp16][n h8][n l8][float s16]                    ammc_q_act_bytes(n)
 *   weight buffer (n = Cout*K, K = taps*Cin, k = tap*Cin + cin): [n fp16][n h8][n l8][Cout floats sum_k|w|][float s16]
 * ammc_pack_nhwc_q: x [b,C,h,w] fp32 NCHW -> q buffer (max|x| reduction + pack).  ammc_pack_conv_weights_q: w
 * [Cout,Cin,3,3] (taps 9) or [Cout,Cin] (taps 1) -> q weight buffer. */
size_t ammc_q_act_bytes(int64_t n);
size_t ammc_q_weight_bytes(int Cout, int K);
int ammc_pack_nhwc_q(const float* x, void* xq, int b, int C, int h, int w, void* stream);
/* the q buffer and the bf16 hi/lo planes [2][b,h,w,C] of the same tensor in one pass (training at precision 2: the forward
 * conv takes the q operand, the weight gradient the bf16 planes); C % 8 == 0 */
int ammc_pack_nhwc_q_planes(const float* x, void* xq, void* xp, int b, int C, int h, int w, void* stream);
int ammc_pack_conv_weights_q(const float* w, void* wq, int Cout, int Cin, int taps, void* stream);
/* forward and data-gradient q weights (w'[ci][co][tap] = w[co][ci][taps-1-tap]; buffer of ammc_q_weight_bytes(Cin, taps*Cout))
 * of one weight tensor in one call: one max|w| reduction serves both */
int ammc_pack_conv_weights_q_pair(const float* w, void* wq, void* wq_dgrad, int Cout, int Cin, int taps, void* stream);
int ammc_conv3x3_bn_relu(const void* xp, const void* wp, const float* scale, const float* shift,
                         void* out_planes, float* out_nchw, const float* res_nchw,
                         int b, int Cin, int Cout, int h, int w, int precision, int relu, void* stream);
/* ---- general layer form of the same engine (the U-Net encoder/decoder around the path,
 *      reference Code/models/unet.py:8-59 double_conv / inconv / down / up, 908-937 UNetMem_v7) --------------------------
 * One descriptor covers: 3x3 conv + folded BN + ReLU (unet.py:11-16), the 1x1 GEMMs, the transposed 2x2 / stride-2
 * conv of `up` (unet.py:46, a 1x1 GEMM with 4*Cout columns whose epilogue scatters column (dy*2+dx)*Cout + co to pixel
 * (2h+dy, 2w+dx)), the final 3x3 conv + bias + tanh (unet.py:918,936), reads from / writes into a channel window of a
 * wider NHWC buffer (so torch.cat([skip, up], 1) at unet.py:58 never copies), and rows wider than 128 pixels. */
typedef struct ammc_conv_layer {
  const void* in_planes;   /* [2][b,h,w,in_cs] bf16 hi/lo planes; the layer reads channels [in_c_off, in_c_off + Cin) */
  int in_cs, in_c_off;     /* in_cs = 0 means Cin */
  const void* wp;          /* [2][Cout][taps*Cin] bf16 planes (ammc_pack_conv_weights* / ammc_pack_convt_weights) */
  int taps;                /* 9 (3x3, padding 1) or 1 */
  const float* scale;      /* [Cout] */
  const float* shift;      /* [Cout]  y = act(acc*scale + shift) */
  int act;                 /* 0 none, 1 ReLU, 2 tanh */
  void* out_planes;        /* [2][b,ho,wo,out_cs] bf16 planes or NULL; channels written at out_c_off */
  int out_cs, out_c_off;   /* out_cs = 0 means dense */
  float* out_nchw;         /* [b,cout_valid,h,w] fp32 or NULL */
  const float* res_nchw;   /* added before both outputs, or NULL */
  int cout_valid;          /* 0 = Cout; < Cout when the weights/scale/shift were zero-padded to a multiple of 64 */
  int b, h, w, Cin, Cout;  /* input feature map; Cin, Cout multiples of 64 */
  int up2x;                /* 1: transposed 2x2 stride-2 conv (taps = 1, Cout = 4*channels, output 2h x 2w) */
  int precision;           /* 3: split-bf16 x3 (fp32 parity); 2: fp16 + e4m3 cross terms (fp32 parity at two pass-
                              equivalents, q operands); 1: single bf16 pass */
  int in_fmt;              /* 0: in_planes / wp are bf16 hi/lo planes; 1: q buffers (required by precision 2) */
  int out_fmt;             /* 0: out_planes are bf16 hi/lo planes; 1: a q buffer.  With precision 2 the layer derives the
                              output scale itself; with precision 1/3 the caller stores it at byte 4*n beforehand */
  int io_bf16;             /* 1: out_nchw / res_nchw point to bf16 NCHW tensors (bf16 feature-I/O variant, BASELINE
                              configs[2]); the sum is formed in fp32 and rounded once at the store.  CTA-pair kernel only
                              (Cout % 256 == 0, act 0/1) */
} ammc_conv_layer;
int ammc_conv_layer_run(const ammc_conv_layer* layer, void* stream);
/* w [Cout,Cin,3,3] (taps=9) or [Cout,Cin] (taps=1) -> wp [2][Cout_pad][taps*Cin_pad], zero-padded rows/columns */
int ammc_pack_conv_weights_padded(const float* w, void* wp, int Cout, int Cin, int Cout_pad, int Cin_pad, int taps,
                                  void* stream);
/* ConvTranspose2d weight [Cin,Cout,2,2] -> wp [2][4*Cout][Cin], row (dy*2+dx)*Cout + co */
int ammc_pack_convt_weights(const float* w, void* wp, int Cin, int Cout, void* stream);
/* x [b,C,h,w] fp32 -> planes [2][b,h,w,C_pad] (channels C..C_pad-1 zero) */
int ammc_pack_nhwc_padded(const float* x, void* xp, int b, int C, int C_pad, int h, int w, void* stream);
/* MaxPool2d(2) (unet.py:33) on NHWC planes: in [2][b,h,w,in_cs] channels [in_c_off, in_c_off+C) -> out [2][b,h/2,w/2,C] */
int ammc_maxpool2_planes(const void* in_planes, int in_cs, int in_c_off, void* out_planes, int b, int h, int w, int C,
                         void* stream);
/* planes [2][b,h,w,cs] channels [c_off, c_off+C) -> fp32 NCHW [b,C,h,w] (hi + lo) */
int ammc_unpack_nhwc(const void* xp, int cs, int c_off, float* x, int b, int C, int h, int w, void* stream);
/* 1 (default): convolutions with Cout % 256 == 0 use the CTA-pair (cta_group::2, M=256) kernel, which at precision 3
 * loads the hi and lo planes of a K block once and issues hi*hi, hi*lo, lo*hi back to back; 3: CTA-pair kernel that
 * streams the K loop three times instead (bit-identical to the single-CTA kernel); 0: always the single-CTA kernel.
 * The variants differ only in fp32 summation order; the switch exists for A/B measurements. */
int ammc_set_conv_pair_mode(int on);
/* 1 (default): 3x3 layers with Cout 64 / 128 on maps >= 64 pixels wide use the halo kernel (one halo tile per 64-channel
 * block, nine shifted tap descriptors over it); 0: always the generic implicit GEMM.  For A/B measurements. */
int ammc_set_conv_halo_mode(int on);
/* 1x1 convolution on the same tensor-core engine (a plain [N,Cin] x [Cout,Cin]^T GEMM, no halo):
 *   wp [2 planes][Cout][Cin] bf16 (pack with ammc_pack_conv_weights_1x1). */
int ammc_pack_conv_weights_1x1(const float* w, void* wp, int Cout, int Cin, void* stream);
int ammc_conv1x1_bn_relu(const void* xp, const void* wp, const float* scale, const float* shift,
                         void* out_planes, float* out_nchw, const float* res_nchw,
                         int b, int Cin, int Cout, int h, int w, int precision, int relu, void* stream);
/* ---- AMFT training path (autograd of double_conv, unet.py:8-20) -------------------------------------------------
 * ammc_bn_batch_stats   training != 0: per-channel batch mean / biased variance of y [b,C,h,w] -> scale = gamma*invstd,
 *                       shift = beta - mean*scale, mean, invstd; running_mean/var updated in place (momentum, unbiased
 *                       variance) like torch.nn.BatchNorm2d.  training == 0: the same quadruple from the running stats.
 *                       workspace: 2*C doubles.
 * ammc_bn_apply         v = relu?(y*scale + shift) written as NHWC bf16 hi/lo planes, NCHW bf16 hi/lo planes and/or fp32
 *                       NCHW (+ res); any subset of the three outputs.
 * ammc_bn_backward      gradient through ReLU + BatchNorm: g [b,C,h,w] wrt the activation -> g_y wrt the conv output as
 *                       NHWC and/or NCHW bf16 planes, g_gamma [C], g_beta [C].  workspace: 2*C doubles.
 * ammc_pack_planes      fp32 -> bf16 hi/lo planes in the same layout ([2][n]).
 * ammc_pack_conv_weights_dgrad   w [Cout,Cin,3,3] -> [2][Cin][9*Cout] (taps flipped): ammc_conv3x3_bn_relu on gradient
 *                       planes with these weights is the data gradient of the convolution.
 * ammc_conv3x3_wgrad    gw [Cout,Cin,3,3] = sum_pixels gy x x(shifted) on tcgen05; operands are the NHWC bf16 planes of
 *                       the output gradient [2][b,h,w,Cout] and of the conv input [2][b,h,w,Cin] (MN-major UMMA).
 * ammc_conv1x1_wgrad    same for a 1x1 conv: gw [Cout,Cin] = gy^T . x, a GEMM whose K axis is the pixels (also used inside
 *                       ammc_mem_bwd for the enc / dec weight gradients). */
int ammc_bn_batch_stats(const float* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
                        float* scale, float* shift, float* mean, float* invstd, void* workspace, size_t workspace_bytes,
                        int b, int C, int h, int w, float momentum, float eps, int training, void* stream);
/* q-format variants for training at precision 2 (the forward and data-gradient convs of `bridge` take fp16 + e4m3 operands):
 * the BatchNorm passes see every value that lands in an operand plane, so the planes' power-of-two scale comes out of their
 * reductions as a bound -- per-channel max|y| in the statistics pass, max|g'| and max|yhat| in the backward reduction.
 *   ammc_bn_batch_stats_q   = ammc_bn_batch_stats (training) + the bound on |act(y*scale+shift)|, left in `workspace`
 *   ammc_bn_apply_q         relu(y*scale+shift) -> bf16 hi/lo NHWC planes (weight-gradient operand, may be NULL) and the q
 *                           buffer out_q (ammc_q_act_bytes); `workspace` is the one ammc_bn_batch_stats_q filled
 *   ammc_bn_backward_q      = ammc_bn_backward, g_y written as bf16 hi/lo planes (may be NULL) and as q planes
 * workspace: ammc_bn_q_workspace_bytes(C) bytes.  Per-rank statistics (no staged / all-reduced form). */
size_t ammc_bn_q_workspace_bytes(int C);
int ammc_bn_batch_stats_q(const float* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
                          float* scale, float* shift, float* mean, float* invstd, void* workspace, size_t workspace_bytes,
                          int b, int C, int h, int w, float momentum, float eps, void* stream);
int ammc_bn_apply_q(const float* y, const float* scale, const float* shift, int relu, void* out_nhwc_planes, void* out_q,
                    const void* workspace, int b, int C, int h, int w, void* stream);
int ammc_bn_backward_q(const float* g, const float* y, const float* scale, const float* shift, const float* mean,
                       const float* invstd, int relu, int training, void* gy_nhwc_planes, void* gy_q, float* g_gamma,
                       float* g_beta, void* workspace, size_t workspace_bytes, int b, int C, int h, int w, void* stream);
int ammc_bn_apply(const float* y, const float* scale, const float* shift, int relu, void* out_nhwc_planes,
                  void* out_nchw_planes, float* out_f32, const float* res, int b, int C, int h, int w, void* stream);
int ammc_bn_backward(const float* g, const float* y, const float* scale, const float* shift, const float* mean,
                     const float* invstd, int relu, int training, void* gy_nhwc_planes, void* gy_nchw_planes,
                     float* g_gamma, float* g_beta, void* workspace, size_t workspace_bytes, int b, int C, int h, int w,
                     void* stream);
/* Staged forms for data-parallel training with GLOBAL-batch BatchNorm statistics (what one GPU running the reference on the
 * whole batch computes): stage 1 leaves this rank's per-channel sums as 2*C doubles at the start of the workspace, the
 * caller all-reduces them (NCCL, SUM), stage 2 finishes with count_total = elements per channel over all ranks.  stage 0 =
 * the single-call forms above.  The backward's stage 1 also writes the LOCAL g_gamma / g_beta (reduced with the other
 * parameter gradients). */
int ammc_bn_batch_stats_staged(const float* y, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, float* scale, float* shift, float* mean, float* invstd, void* workspace,
                               size_t workspace_bytes, int b, int C, int h, int w, float momentum, float eps, int training,
                               int stage, double count_total, void* stream);
int ammc_bn_backward_staged(const float* g, const float* y, const float* scale, const float* shift, const float* mean,
                            const float* invstd, int relu, int training, void* gy_nhwc_planes, void* gy_nchw_planes,
                            float* g_gamma, float* g_beta, void* workspace, size_t workspace_bytes, int b, int C, int h, int w,
                            int stage, double count_total, void* stream);
int ammc_pack_planes(const float* x, void* xp, int64_t n, void* stream);
int ammc_pack_conv_weights_dgrad(const float* w, void* wp, int Cout, int Cin, void* stream);
int ammc_conv3x3_wgrad(const void* gy_nhwc_planes, const void* x_nhwc_planes, float* gw, int b, int Cin, int Cout,
                       int h, int w, int precision, void* stream);
/* BatchNorm (eval) folding: scale/shift [C] from gamma, beta, running_mean, running_var. */
int ammc_conv1x1_wgrad(const void* gy_nhwc_planes, const void* x_nhwc_planes, float* gw, int b, int Cin, int Cout, int h,
                       int w, int precision, void* stream);
int ammc_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                 float* scale, float* shift, int C, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Scoring.  ammc_psnr_batch replaces the per-frame psnr_error calls of the scoring loop
 * (reference Code/utils/utils.py:130-148 called at Code/run_helper/test_helper.py:445-452):
 *   psnr[i] = 10*log10( 1 / ( (1/elems) * sum_j ((gt+1)/2 - (gen+1)/2)^2 ) ),  gen/gt [n, elems]
 * ammc_score_reduce replaces norm_score + mixing + smoothing (Code/main/eval_metric.py:405-427):
 *   img/fea [T_total] concatenated per-video records, offsets [V+1] int64 (video v = [offsets[v], offsets[v+1]))
 *   scores [T_total - 4*V]; one_minus_lam* are the fp32 roundings of the host doubles (1-lam), as numpy uses.
 *   Bit-exact with the numpy float32 reference (no FMA contraction, IEEE division).
 * ------------------------------------------------------------------------------------------------- */
size_t ammc_psnr_workspace_bytes(int n, int64_t elems);
int ammc_psnr_batch(const float* gen, const float* gt, float* psnr, void* workspace, size_t workspace_bytes,
                    int n, int64_t elems, void* stream);

size_t ammc_score_workspace_bytes(int64_t t_total, int n_videos);
int ammc_score_reduce(const float* img, const float* fea, const int64_t* offsets, int n_videos,
                      float one_minus_lam1, float lam1, float one_minus_lam2, float lam2,
                      float* scores, void* workspace, size_t workspace_bytes, int64_t t_total, void* stream);

/* ROC-AUC of scores against labels (labels[i] == pos_label marks the positive class), the value
 * sklearn.metrics.auc(roc_curve(labels, scores, pos_label)) returns at reference Code/main/eval_metric.py:428-429:
 * bitonic sort + tie-aware rank sum, exact in integers, final ratio in fp64.  auc: device double[1]. */
size_t ammc_auc_workspace_bytes(int64_t T);
int ammc_roc_auc(const float* scores, const int8_t* labels, int pos_label, double* auc, void* workspace,
                 size_t workspace_bytes, int64_t T, void* stream);

/* ---- frame / flow preprocessing in front of the generator --------------------------------
 * Replaces the per-frame CPU work of Code/dataset/two_stream_dataset.py:72-99 after decoding (cv2.cvtColor + cv2.resize +
 * ToTensor + Normalize, resp. cv2.resize + the flow scaling), bit-exactly:
 *   ammc_preprocess_frames_u8  decoded BGR uint8 [n,h0,w0,3] -> fp32 RGB [n,3,H,W] in (-1,1)
 *   ammc_preprocess_flow       .flo payload fp32 [n,h0,w0,2]  -> fp32 [n,2,H,W]; ch0 = resized u / H, ch1 = ch0 / W (the
 *                              loader derives channel 1 from channel 0, two_stream_dataset.py:94-95)
 * The host uploads uint8 frames (a quarter of the bytes of the fp32 tensors). */
int ammc_preprocess_frames_u8(const uint8_t* frames_bgr, float* out, int n, int h0, int w0, int H, int W, void* stream);
int ammc_preprocess_flow(const float* flow, float* out, int n, int h0, int w0, int H, int W, void* stream);
/* bf16 -> fp32 widening of a host-boundary buffer (bf16 feature I/O, BASELINE configs[2]): exact, the fp32-parity path
 * then runs unchanged on the widened values. */
int ammc_cast_bf16_f32(const void* src_bf16, float* dst, int64_t n, void* stream);

/* ---- image-space generator losses (Code/models/losses/losses_utils.py:17-59,124-129) --------
 * out2[0] = Intensity_Loss(l_num=2): mean over (b,h,w) of the channel-wise L2 norm of gen - gt;
 * out2[1] = Gradient_Loss(alpha=1):  mean of |dx| + |dy| of the channel-summed difference (zero padding left / top).
 * bwd: grad_gen = g_int * d out2[0]/d gen + g_gd * d out2[1]/d gen; g_int / g_gd are device scalars (either may be NULL). */
size_t ammc_frame_losses_workspace_bytes(int n, int H, int W);
int ammc_frame_losses_fwd(const float* gen, const float* gt, float* out2, void* workspace, size_t workspace_bytes, int n, int C,
                          int H, int W, void* stream);
int ammc_frame_losses_bwd(const float* gen, const float* gt, const float* g_int, const float* g_gd, float* grad_gen, int n,
                          int C, int H, int W, void* stream);

/* ---- element-wise mean objectives of the training step (Code/models/losses/losses_utils.py:10-15,103-113; consumers
 * Code/models/losses/loss_zoo.py:331-336, Code/run_helper/train_helper.py:318-326) ---------------------------------------
 *   AMMC_OBJ_L1       Flow_Loss          out1[0] = mean |a - b|
 *   AMMC_OBJ_LSGAN_G  Adversarial_Loss   out1[0] = mean (a - 1)^2 / 2                    (b unused, may be NULL)
 *   AMMC_OBJ_LSGAN_D  Discriminate_Loss  out1[0] = mean (a - 1)^2 / 2 + mean b^2 / 2     (a = real map, b = fake map)
 * a and b hold n fp32 elements each.  bwd: grad_a = g * d out1/d a, grad_b = g * d out1/d b; g is a device scalar, either
 * gradient pointer may be NULL.  One pass + a fixed-order final sum (deterministic); needs workspace_bytes(n). */
#define AMMC_OBJ_L1 0
#define AMMC_OBJ_LSGAN_G 1
#define AMMC_OBJ_LSGAN_D 2
size_t ammc_elem_loss_workspace_bytes(int64_t n);
int ammc_elem_loss_fwd(const float* a, const float* b, float* out1, int mode, int64_t n, void* workspace,
                       size_t workspace_bytes, void* stream);
int ammc_elem_loss_bwd(const float* a, const float* b, const float* g, float* grad_a, float* grad_b, int mode, int64_t n,
                       void* stream);

/* ---- generator objective of the joint training step in one call (Twostream_vq_Loss.forward,
 * Code/models/losses/loss_zoo.py:312-350; ctor base_Loss, loss_zoo.py:15-45) ----------------------------------------------
 *   out8[0] = lam_adv * adv + lam_gdl * gd + lam_flow * flow + lam_lp * int + lam_latent * latent + lam_lp_op * int_op
 *   out8[1..6] = adv (Adversarial_Loss of d_gen), flow (Flow_Loss), int / gd (Intensity / Gradient loss of the rgb pair),
 *   int_op (Intensity loss of the flow-stream pair), latent (sum of the n_latent commit-loss values); out8[7] = 0.
 * rgb pair [n_rgb, C_rgb, H_rgb, W_rgb], op pair [n_op, C_op, H_op, W_op], flows n_flow elements each, d_gen n_dgen elements.
 * fwd: four partial-sum passes + one final kernel (fixed summation order).  bwd: g8 = gradient w.r.t. out8 (device, 8 floats);
 * scal8 (device, 8 floats, output) receives the chain-rule scalars [g8[0], d/d adv, d/d flow, d/d int, d/d gd, d/d int_op,
 * d/d latent, 0]; each non-NULL gradient pointer is written by one pass (grad of the latent values = scal8[6] each). */
size_t ammc_gen_objective_workspace_bytes(int n_rgb, int H_rgb, int W_rgb, int n_op, int H_op, int W_op, int64_t n_flow,
                                          int64_t n_dgen);
int ammc_gen_objective_fwd(const float* rgb_out, const float* rgb_tgt, const float* op_out, const float* op_tgt,
                           const float* flow_pred, const float* flow_gt, const float* d_gen, const float* latent, int n_rgb,
                           int C_rgb, int H_rgb, int W_rgb, int n_op, int C_op, int H_op, int W_op, int64_t n_flow, int64_t n_dgen,
                           int n_latent, float lam_adv, float lam_gdl, float lam_flow, float lam_lp, float lam_latent,
                           float lam_lp_op, float* out8, void* workspace, size_t workspace_bytes, void* stream);
int ammc_gen_objective_bwd(const float* rgb_out, const float* rgb_tgt, const float* op_out, const float* op_tgt,
                           const float* flow_pred, const float* flow_gt, const float* d_gen, const float* g8, int n_rgb, int C_rgb,
                           int H_rgb, int W_rgb, int n_op, int C_op, int H_op, int W_op, int64_t n_flow, int64_t n_dgen,
                           float lam_adv, float lam_gdl, float lam_flow, float lam_lp, float lam_latent, float lam_lp_op,
                           float* scal8, float* grad_rgb, float* grad_op, float* grad_flow_pred, float* grad_d_gen, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AMMC_B200_H_ */
