/*
 * tqdne_b200.h -- C-ABI of the B200-native EDM sampling engine (libtqdne_b200.so).
 *
 * The reference (highfem/tqdne) has no FFI layer: its "operator API" is the Python module
 * surface.  This header is the boundary a maintainer binds from Python with
 * ctypes (INTEGRATION.md shows the stub); every entry point cites the reference call it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers are raw CUDA addresses owned by the caller
 *     (PyTorch in the shipped host code); the library never allocates activation memory
 *   - every function returns 0 on success, non-zero on failure; tq_last_error() describes it
 *   - `stream` is a cudaStream_t passed as void*
 *   - activations are channels-last: [N, H, W, C] (2D) or [N, 1, L, C] (1D), dense
 *   - `dtype`: TQ_BF16 -> tcgen05/TMEM/TMA tensor path, TQ_F32 -> FFMA parity path
 *   - an op is appended to a `tq_plan` once (tensor maps are encoded then) and replayed with
 *     tq_plan_run(); a plan is a straight-line list of kernel launches, CUDA-graph capturable
 */
#ifndef TQDNE_B200_H
#define TQDNE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ABI history of round 2 (tq_abi_version() must equal the header the caller was built against):
 *   13  deterministic GroupNorm statistics: tq_conv_desc.stats_parts, tq_gn_desc.parts0/1, tq_conv_stats_parts,
 *       tq_groupnorm_ws_floats                      14  t_next / t_value on the sampler kernels
 *   15  tq_gn_bwd_desc.dbias0/1                      16  tq_repack_batch_prepare / tq_repack_batch_run
 *   17  tq_gn_desc.film (FiLM ResBlocks)             18  tq_attn_desc.causal
 *   19  tq_plan_add_resample2                        20  tq_attn_desc.lse, tq_attention_writes_lse,
 *                                                        tq_attention_backward(..., lse_given, stream)            */
#define TQ_ABI_VERSION 20

enum { TQ_BF16 = 0, TQ_F32 = 1, TQ_F64 = 2 };

/* ---- library ------------------------------------------------------------------------------ */
int         tq_abi_version(void);
const char* tq_last_error(void);
/* number of kernels launched by this library in this process since load / since last reset */
int64_t     tq_launch_count(void);
void        tq_launch_count_reset(void);

/* ---- plan (straight-line launch list) ------------------------------------------------------ */
typedef struct tq_plan tq_plan;
tq_plan* tq_plan_create(void);
void     tq_plan_destroy(tq_plan* p);
int      tq_plan_num_ops(const tq_plan* p);
/* launch ops [first, last) on `stream`; last < 0 means "to the end" */
int      tq_plan_run(tq_plan* p, void* stream);
int      tq_plan_run_range(tq_plan* p, int first, int last, void* stream);
/* capture the whole plan into a CUDA graph once; later tq_plan_run() replays the graph */
int      tq_plan_enable_graph(tq_plan* p, int enable);
/* cudaMemsetAsync(ptr, 0, bytes) as a plan op; at_front != 0 inserts it before every op added so
 * far (the statistics arena of a plan is cleared once, ahead of the first producer)              */
int      tq_plan_add_memset(tq_plan* p, void* ptr, int64_t bytes, int32_t at_front);
/* name of the kernel behind op i (for profiles and tests) */
const char* tq_plan_op_name(const tq_plan* p, int i);

/* ---- implicit-GEMM convolution / linear  -------------------------------------------------- *
 * Replaces: nn.Conv1d/Conv2d built by conv_nd (tqdne/nn.py:16-24) at every call site of the
 * denoiser and decoder (tqdne/unet.py:88,102,108-112,233,357; tqdne/blocks.py:56,94,128,134,
 * 241,249,253,316,348,387,430), the nearest-upsample + conv pair (blocks.py:59-65), the skip
 * concat feeding a conv (unet.py:396) and nn.Linear of emb_layers (unet.py:93-99).
 *
 * The reduction is described as a list of 64-channel K-slices.  Slice s multiplies the input box
 * of source `src` shifted by (dx, dy) at channels [c0, c0+64) with weight columns
 * [kb*64, kb*64+64).  A 3x3 "same" conv is 9*Cin/64 slices; a stride-2 conv reads four parity
 * views of its input; a fused nearest-upsample conv runs 4 output-parity classes over the
 * low-resolution grid; a skip concat or a fused 1x1 shortcut is just more slices / sources.    */
typedef struct {
    const void* ptr;      /* element (n=0,y=0,x=0,c=0) of the view                              */
    int32_t N, H, W, C;   /* extent of the view (OOB reads are zero)                            */
    int64_t sn, sy, sx;   /* element strides of n, y, x (channel stride is 1)                   */
} tq_src;

typedef struct {
    int16_t src, dx, dy, rsv;
    int32_t c0;           /* first channel of the slice inside the source                        */
    int32_t kb;           /* weight K block (64 columns each)                                    */
} tq_slice;

typedef struct {
    int32_t dtype;            /* TQ_BF16 | TQ_F32 : type of sources, weights, residual           */
    int32_t N, H, W;          /* grid the 128-row tiles walk (output grid of one parity class)   */
    int32_t cout;             /* real output channels                                            */
    int32_t cout_pad;         /* rows of the weight matrix (multiple of 64)                      */
    int32_t ktot;             /* columns of the weight matrix (multiple of 64)                   */
    int32_t num_srcs;         /* 1..4                                                            */
    int32_t num_classes;      /* 1, or 4 for the fused nearest-upsample conv                     */
    int32_t num_slices;       /* slices per class                                                */
    tq_src  srcs[4];
    const tq_slice* slices;   /* HOST array [num_classes][num_slices], copied by the call        */
    const void*  weights;     /* device [cout_pad][ktot], K contiguous                           */
    const float* bias;        /* device [cout_pad] or NULL                                       */
    const float* emb;         /* device: per-sample channel add emb[n*emb_ld + c] or NULL        */
    int32_t emb_ld;
    const void* residual;     /* device, same layout as out, dtype = `dtype`, or NULL            */
    void*   out;
    int32_t out_dtype;        /* TQ_BF16 | TQ_F32                                                */
    int64_t out_sn, out_sy, out_sx;   /* element strides of the output                           */
    int64_t out_class_off[4];         /* element offset of each parity class                     */
    int32_t block_n;          /* 0 = auto, else 64 / 128 / 256                                   */
    float*  stats;            /* device [N][stats_parts][cout][2] fp32 or NULL: (sum, sum of squares) of  *
                               * the STORED output per sample, PART and channel, written in the epilogue  *
                               * with plain stores -- a part is one output tile of the sample, every slot  *
                               * has exactly one writer, nothing is accumulated with atomics.  Feeds      *
                               * tq_gn_desc.stats0/1 of the consuming GroupNorm, which adds the parts in   *
                               * index order: the normalisation never re-reads the tensor for its         *
                               * statistics and the result is bit-reproducible.  Allocate it zero-filled   *
                               * (slots a geometry never writes stay zero).  bf16 outputs only on the      *
                               * tensor path.                                                              */
    int32_t cta_group;        /* 0 = auto, 1 = one CTA per 128-row tile, 2 = CTA pair per 256-row   *
                               * tile (tcgen05.mma.cta_group::2)                                    */
    int32_t stats_parts;      /* parts dimension of `stats`: must equal tq_conv_stats_parts(d)      */
} tq_conv_desc;
int tq_plan_add_conv(tq_plan* p, const tq_conv_desc* d);
/* number of statistics parts per sample the kernel chosen for this geometry (dtype, H, W, num_classes) writes */
int32_t tq_conv_stats_parts(const tq_conv_desc* d);

/* ---- GroupNorm(32) [+ SiLU] over a (virtual) channel concat -------------------------------- *
 * Replaces: GroupNorm32 / normalization (tqdne/nn.py:11-13,90-105) + nn.SiLU in in_layers /
 * out_layers (unet.py:85-103, blocks.py:238-250), attention norm (blocks.py:126), `out`
 * (unet.py:354-356), fed by th.cat([h, hs.pop()], dim=1) (unet.py:396) without materialising it.
 * x0:[N,P,C0] (+ x1:[N,P,C1]) -> y:[N,P,C0+C1]; stats in fp32; eps as given; 32 groups.
 * stats0 / stats1 ([N][parts0][C0][2], [N][parts1][C1][2], written by the producing conv's       *
 * epilogue, tq_conv_desc.stats) skip the statistics pass, leaving ONE streaming pass (2 B read  *
 * + 2 B written per element in bf16).  Both or neither must be given for a two-source norm.     *
 * `ws` is a caller-provided fp32 scratch of tq_groupnorm_ws_floats(d) floats (0 for most fused  *
 * norms): the stand-alone statistics pass and, for tensors cut into many parts, the per-sample  *
 * group reduction live there.  No atomics anywhere: results are bit-reproducible.               */
typedef struct {
    int32_t dtype;         /* TQ_BF16 | TQ_F32 for x0, x1, y                                      */
    int32_t N, P, C0, C1;  /* P = H*W positions                                                   */
    const void* x0; const void* x1;
    const float* gamma; const float* beta;   /* [C0+C1]                                           */
    float   eps;
    int32_t silu;          /* apply x*sigmoid(x) after the affine                                  */
    void*   y;
    float*  ws;
    const float* stats0; const float* stats1;
    int32_t parts0, parts1; /* parts dimension of stats0 / stats1 (tq_conv_desc.stats_parts of the producers) */
    /* training only (bf16): y = dropout(act(GroupNorm(x))), nn.Dropout of ResBlock.out_layers (tqdne/unet.py:100-108).
     * The decision of element i is hash(*drop_seed + site, i) < drop_p; NULL / 0 = no dropout (every sampling plan).   */
    const uint64_t* drop_seed; float drop_p; int32_t drop_site;
    /* FiLM, ResBlock(use_scale_shift_norm=True) (tqdne/unet.py:135-139): y = act(GroupNorm(x) * (1 + scale) + shift) with
     * scale[n][c] = film[n * film_ld + c], shift[n][c] = film[n * film_ld + C0 + C1 + c] (fp32; the two halves of the
     * block's embedding projection, th.chunk(emb_out, 2, dim=1)).  NULL = plain GroupNorm.                            */
    const float* film; int32_t film_ld;
} tq_gn_desc;
int tq_plan_add_groupnorm(tq_plan* p, const tq_gn_desc* d);
/* fp32 scratch floats `ws` must hold for this op on the current device (0 = none needed), < 0 on bad arguments */
int64_t tq_groupnorm_ws_floats(const tq_gn_desc* d);

/* ---- attention core -------------------------------------------------------------------------- *
 * Replaces: QKVAttention.forward (tqdne/blocks.py:156-190).  qkv:[N,T,3*heads*d] channels-last
 * with channel = third*(heads*d) + head*d + c, out:[N,T,heads*d];
 * w = softmax_fp32((q*s)^T (k*s)), s = d^-1/4, out = w v.
 * `causal` != 0: key s > query t is masked out before the softmax (use_causal_mask, blocks.py:181-186; FFMA kernels). */
typedef struct {
    int32_t dtype; int32_t N, T, heads, d;
    const void* qkv; void* out;
    int32_t causal;
    /* optional (training): [N][heads][T] fp32, L_i = log2 sum_j 2^(s_ij d^-1/2 log2 e) of every query row, the statistic
     * tq_attention_backward otherwise recomputes.  Written by the multi-block tensor-core kernel (bf16, d in {64, 128},
     * 128 < T <= 512); tq_attention_writes_lse(d) tells whether the kernel chosen for `d` does.                         */
    float* lse;
} tq_attn_desc;
int32_t tq_attention_writes_lse(const tq_attn_desc* d);
int tq_plan_add_attention(tq_plan* p, const tq_attn_desc* d);

/* ---- small dense layers in fp32 (embedding MLPs) ---------------------------------------------- *
 * Replaces: nn.Linear / nn.SiLU of time_mlp, cond_mlp (tqdne/unet.py:210-227), their sum
 * (unet.py:383-388) and the SiLU in front of every emb_layers Linear (unet.py:92).
 *   v[m][j]      = sum_k act_in(x[m or 0][k]) * W[j][k] + b[j] (+ add[m or 0][j])
 *   y[m][j]      = v                          (fp32, optional)
 *   y_act[m][j]  = SiLU(v)                    (optional, stored as y_act_dtype: TQ_F32 | TQ_BF16)
 * act_in: 0 none, 1 SiLU.  x_rows / add_rows == 1 broadcasts a single row to all M outputs.        */
typedef struct {
    int32_t M, K, Nout, x_rows;
    const float* x; const float* W; const float* b;
    int32_t act_in;
    const float* add; int32_t add_rows;
    float* y;
    void*  y_act; int32_t y_act_dtype;
} tq_linear_desc;
int tq_plan_add_linear(tq_plan* p, const tq_linear_desc* d);
/* feat[M, 2*half] = [sin(2*pi*t*W), cos(2*pi*t*W)],  t:[M] read from device                      */
int tq_plan_add_fourier(tq_plan* p, const float* t, const float* W, int32_t M, int32_t half, float* feat);
/* The conv-less resamplers of conv_resample=False models, channels-last x:[N,H,W,C] (H = 1 for 1-D), C % 8 == 0:
 * mode 0 = nn.AvgPool{1,2}d(kernel 2, stride 2) (tqdne/blocks.py:104) -> y:[N,H/2,W/2,C] (odd sizes floor, like torch);
 * mode 1 = F.interpolate(scale_factor=2, mode="nearest") (blocks.py:59-62) -> y:[N,2H,2W,C].                        */
int tq_plan_add_resample2(tq_plan* p, int32_t dtype, const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C,
                          int32_t mode);
/* y[N, C] = mean over the P positions of x[N, P, ld] (fp32 channels-last, first C channels).
 * Replaces: th.mean(h, dim=spatial) in LithningClassifier.embed (tqdne/classifier.py:51-53).      */
int tq_plan_add_spatial_mean(tq_plan* p, const float* x, int32_t N, int32_t P, int32_t C, int32_t ld, float* y);

/* ---- convolution weight gradient (first kernel of the training-step row) ---- *
 * Replaces: the weight / bias gradient autograd computes for nn.Conv1d (tqdne/nn.py:16-24, stride 1,
 * padding "same") inside LightningEDM.step (tqdne/edm.py:115-134).
 *   x  : [N, L, cin]  bf16 channels-last (the layer input that the forward plan kept)
 *   dy : [N, L, cout] bf16 channels-last (gradient of the layer output)
 *   dw : [cout, taps, cin] fp32, db : [cout] fp32 or NULL -- ACCUMULATED INTO (zero them first):
 *        dw[co][t][ci] += sum_{n,l} dy[n][l][co] * x[n][l + t - taps/2][ci],  db[co] += sum_{n,l} dy[n][l][co]
 * cin and cout must be multiples of 64, taps odd and <= 7.  A layer whose input is a channel concat calls this once
 * per source: dw_ld = input channels of the whole layer (0 = cin), ci_off = this source's first channel.   */
int tq_conv1d_wgrad(const void* x, const void* dy, float* dw, float* db, int32_t N, int64_t L, int32_t cin,
                    int32_t cout, int32_t taps, int32_t dw_ld, int32_t ci_off, void* stream);
/* 2-D counterpart (groundwork for training the 2-D UNets / autoencoder; nn.Conv2d, stride 1, padding "same"):
 *   x [N,H,W,cin], dy [N,H,W,cout] bf16 channels-last; dw [cout, kh*kw, cin] fp32 ACCUMULATED INTO:
 *   dw[co][ky*kw+kx][ci] += sum_{n,y,x} dy[n][y][x][co] * x[n][y+ky-kh/2][x+kx-kw/2][ci]; the bias gradient is not computed. */
int tq_conv2d_wgrad(const void* x, const void* dy, float* dw, int32_t N, int32_t H, int32_t W, int32_t cin, int32_t cout,
                    int32_t kh, int32_t kw, void* stream);
/* out[n*out_ld + c] += sum_p dy[n][p][c] (dy bf16 [N,P,C], out fp32 rows of length out_ld, 0 = C): gradient of the per-sample embedding
 * term a ResBlock adds after its first convolution (tqdne/unet.py:129-141).                          */
int tq_sample_channel_sums(const void* dy, float* out, int32_t out_ld, int32_t N, int64_t P, int32_t C, void* stream);

/* ---- GroupNorm(32) [+ SiLU] backward (training-step row) --------------------- *
 * Replaces: autograd through GroupNorm32 + nn.SiLU (tqdne/nn.py:11-13,90-105, tqdne/unet.py:85-88,100-103).
 * Same tensors as tq_gn_desc (virtual concat of x0 [N,P,C0] and x1 [N,P,C1], stats0/stats1 = the forward
 * per-(sample, channel) sum / sum of squares), plus dy [N,P,C0+C1]; writes dx0 / dx1 and ACCUMULATES
 * dgamma / dbeta [C0+C1] (either may be NULL).  ws: [N][C0+C1][2] fp32 scratch (cleared by the call).   */
typedef struct {
    int32_t dtype; int32_t N, P, C0, C1;
    const void* x0; const void* x1; const void* dy;
    const float* gamma; const float* beta;
    float eps; int32_t silu;
    const float* stats0; const float* stats1;   /* [N][parts0][C0][2], [N][parts1][C1][2] (tq_conv_desc.stats) */
    int32_t parts0, parts1;
    float* ws;
    void* dx0; void* dx1;
    float* dgamma; float* dbeta;
    const void* dx_add0; const void* dx_add1;   /* optional: gradient of x0 / x1 from their other consumer, added to dx */
    float* dx_sum; int32_t dx_sum_ld;            /* optional: dx_sum[n*ld + c] += sum_p dx0[n][p][c] (embedding gradient) */
    const uint64_t* drop_seed; float drop_p; int32_t drop_site;   /* the forward's fused dropout (tq_gn_desc), or NULL */
    float* dbias0; float* dbias1;   /* optional: dbias[c] += sum over samples and positions of dx0 / dx1 -- the bias gradient of
                                     * the convolution that produced x0 / x1 when dx (with dx_add) is that tensor's WHOLE gradient */
} tq_gn_bwd_desc;
int tq_gn_silu_backward(const tq_gn_bwd_desc* d, void* stream);

/* ---- attention core backward (training-step row) -------------------------------- *
 * Replaces: autograd through QKVAttention.forward (tqdne/blocks.py:156-190).  Same layouts as tq_attn_desc:
 * qkv [N,T,3*heads*d] (forward input), out [N,T,heads*d] (forward output), dout = gradient of out, all bf16;
 * writes dqkv [N,T,3*heads*d] bf16.  ws: 2*N*heads*T floats of scratch (row log-sum-exp and D_i).
 * `lse_given` != 0: ws[0 .. N*heads*T) already holds the log-sum-exp the forward wrote (tq_attn_desc.lse = ws) and the
 * kernel skips recomputing it (one S = Q K^T over all keys and two softmax passes per query block).
 * tcgen05 kernels; head dim 64, 32 < T <= 512 (the 1D UNet's attention blocks).                          */
int tq_attention_backward(const void* qkv, const void* out, const void* dout, void* dqkv, float* ws, int32_t N,
                          int32_t T, int32_t heads, int32_t d, int32_t lse_given, void* stream);

/* ---- small kernels of the training step ------------------------------------------ *
 * tq_rows_op: rows of a channels-last bf16 tensor [N, L, C]; L_dst = rows of dst per sample.
 *   mode 0 zero_stuff (dst[2j] = src[j], dst[2j+1] = 0: dY of a stride-2 conv on the stride-1 grid, blocks.py:93-101),
 *   mode 1 nearest x2 upsample (blocks.py:59-65), mode 2 pair_sum (its gradient), mode 3 dst = src * aux (dropout mask),
 *   mode 4 dst = src + aux (two gradient paths of one tensor).
 * tq_linear_backward: v = act_in(x) W^T + b in fp32 (time / cond MLPs, emb_layers; unet.py:210-227,92):
 *   dW += dy^T act_in(x), db += sum dy, dx = act_in'(x) * (dy W); act_in 0 none / 1 SiLU; dx, dW, db may be NULL.
 * tq_edm_noise / tq_edm_loss: LightningEDM.step (edm.py:115-134): xn = y + sigma*noise, network input bf16(c_in xn)
 *   with padded channels; loss = mean(w (c_out F + c_skip xn - y)^2) and dF (bf16, padded) = its gradient wrt F.
 * tq_dropout_mask: 0 / 1/(1-p) scale tensor (nn.Dropout, unet.py:100-108).
 * tq_adam_ema_step: torch.optim.Adam defaults (edm.py:240-251) + EMA lerp (ema.py:24-28) over flat fp32 arrays.    */
int tq_rows_op(const void* src, const void* aux, void* dst, int32_t mode, int64_t N, int64_t L_dst, int32_t C, void* stream);
int tq_linear_backward(const float* dy, const float* x, const float* W, int32_t act_in, float* dx, float* dW, float* db, int32_t M,
                       int32_t K, int32_t Nout, void* stream);
int tq_edm_noise(const float* y, const float* noise, const float* sigma, float* xn, void* xin, int64_t N, int64_t P, int32_t C,
                 int32_t Cpad, float sigma_data, void* stream);
int tq_edm_loss(const float* F, int32_t Cf, const float* xn, const float* y, const float* sigma, void* dF, float* loss, int64_t N,
                int64_t P, int32_t C, int32_t Cpad, float sigma_data, void* stream);
int tq_dropout_mask(void* mask, int64_t n, uint64_t seed, float p, void* stream);
/* dst = src * the same scale, generated on the fly (forward and backward call it with the same seed): bf16, n % 8 == 0 */
int tq_dropout_apply(const void* src, void* dst, int64_t n, uint64_t seed, float p, void* stream);
/* bf16 operand copies of one convolution from its fp32 master [Op][k][Ip]: fwd [Op][k*Ip] (plain cast) and / or
 * bwd [Cs][k*Op] = master[co][k-1-t][ci_off+ci] (input-gradient operand: taps flipped, in / out transposed).      */
int tq_repack_conv_weights(const float* master, void* fwd, void* bwd, int32_t Op, int32_t k, int32_t Ip, int32_t ci_off,
                           int32_t Cs, void* stream);
/* All operand copies of a model in ONE launch (the per-convolution calls above are ~170 launches of 2-5 us each, a
 * serial chain even inside a CUDA graph: 0.68 ms of a 14.7 ms training step).  Fill `jobs` (pointers = device addresses,
 * one job per copy: exactly one of fwd / bwd set), call tq_repack_batch_prepare on the HOST array -- it assigns every job
 * its range of thread blocks and returns their total (-1 on a bad job) -- copy the array to device memory and launch
 * tq_repack_batch_run with that device copy after every optimiser step.                                               */
typedef struct {
    const float* master;   /* fp32 [Op][k][Ip] */
    void* fwd;             /* bf16 [Op][k*Ip] or NULL */
    void* bwd;             /* bf16 [Cs][k*Op] or NULL */
    int32_t Op, k, Ip, ci_off, Cs;
    int32_t block0;        /* filled by tq_repack_batch_prepare: first thread block of this job */
    int32_t nblocks;       /*                                     and how many it owns          */
    int32_t pad_;
} tq_repack_job;
int64_t tq_repack_batch_prepare(tq_repack_job* jobs_host, int32_t n_jobs);
int     tq_repack_batch_run(const tq_repack_job* jobs_device, int32_t n_jobs, int64_t total_blocks, void* stream);
int tq_adam_ema_step(float* param, const float* grad, float* m, float* v, float* ema, int64_t n, float lr, float beta1, float beta2,
                     float eps, int64_t step, float ema_decay, float grad_scale, void* stream);

/* ---- sampler element-wise steps ------------------------------------------------------------------ *
 * Replaces: LightningEDM.forward pre/post scaling (tqdne/edm.py:105-113) and the Heun/Euler
 * state update of sample_deterministically / sample_stochastically (edm.py:171-230).
 * State x is fp64 channels-last [N,P,C]; the denoiser input is written with C padded to Cpad.
 * All scalars are passed by value per call (they are not part of a plan).
 *
 *   tq_edm_precondition : xin = dtype( float(x) * c_in )            (pad channels = 0)
 *   tq_edm_euler        : D = F*c_out + c_skip*float(x);  d = (x - D)/sigma;  x1 = x + d*dt
 *                         optionally xin = dtype(float(x1) * c_in_next)
 *   tq_edm_heun         : D' = F*c_out' + c_skip'*float(x1); d' = (x1 - D')/sigma';
 *                         x = x + dt*(0.5*d + 0.5*d');  optionally xin = dtype(float(x)*c_in_next)
 * F is the fp32 channels-last network output [N,P,Cf] (row stride Cf >= C).
 * t_next / t_value (t_next may be NULL): also store the NEXT denoiser call's noise-conditioning scalar
 * c_noise = ln(sigma)/4 (edm.py:36-37) into the plan's time input -- a 4-byte cudaMemcpyAsync between two graph
 * replays cost 0.22 ms per call (copy-engine hand-over), a store from the kernel that runs there anyway is free.   */
int tq_edm_precondition(const double* x, void* xin, int32_t dtype, int64_t NP, int32_t C, int32_t Cpad,
                        float c_in, float* t_next, float t_value, void* stream);
int tq_edm_euler(const double* x, const float* F, int32_t Cf, double* d, double* x1,
                 void* xin, int32_t dtype, int64_t NP, int32_t C, int32_t Cpad,
                 float c_out, float c_skip, float sigma, float dt, float c_in_next, int32_t write_xin,
                 float* t_next, float t_value, void* stream);
int tq_edm_heun(double* x, const double* x1, const double* d, const float* F, int32_t Cf,
                void* xin, int32_t dtype, int64_t NP, int32_t C, int32_t Cpad,
                float c_out, float c_skip, float sigma_next, float dt, float c_in_next, int32_t write_xin,
                float* t_next, float t_value, void* stream);
/* x += noise * scale  (stochastic sampler churn, edm.py:205-207), fp64 */
int tq_edm_add_noise(double* x, const double* noise, double scale, int64_t n, void* stream);

/* ---- layout ------------------------------------------------------------------------------------- *
 * [N,C,P] (reference NCHW/NCL, dtype_in) <-> [N,P,Cpad] channels-last (dtype_out), pad = 0.         */
int tq_nchw_to_nhwc(const void* src, int32_t dtype_in, void* dst, int32_t dtype_out,
                    int32_t N, int32_t C, int64_t P, int32_t Cpad, void* stream);
int tq_nhwc_to_nchw(const void* src, int32_t dtype_in, int32_t Cld, void* dst, int32_t dtype_out,
                    int32_t N, int32_t C, int64_t P, void* stream);

/* ---- representation inverses ----------------------------------------------------------------------- *
 * tq_logspec_griffinlim replaces LogSpectrogram.invert_representation (tqdne/representation.py:
 * 152-175) with librosa 0.11 griffinlim(n_iter, hop, n_fft, random_state=0) semantics:
 * rep:[items, n_fft/2, frames] in [-1,1] (reference NCHW order, any float dtype cast to fp32 by the
 * caller) -> wave:[items, hop*(frames-1)] (float for TQ_F32, double for TQ_F64).
 * phase0:[n_fft/2+1, frames, 2] float64 are the unit phasors (cos, sin) of the initial phase angles
 * 2*pi*RandomState(0).random(), shared by all items.  ws: scratch, size from
 * tq_griffinlim_ws_bytes().  precision: TQ_F64 (the reference under its locked NumPy 2 runs
 * complex128; the parity mode and the default of the Python front-end) or TQ_F32 (fast mode).
 * All iterations run in ONE launch; the result is bit-reproducible (no atomics).                       */
int64_t tq_griffinlim_ws_bytes(int32_t items, int32_t n_fft, int32_t frames, int32_t precision);
int tq_logspec_griffinlim(const float* rep, const double* phase0, void* wave, int32_t items,
                          int32_t n_fft, int32_t hop, int32_t frames, int32_t n_iter,
                          double log_clip, double log_max, double momentum, int32_t precision,
                          void* ws, void* stream);
/* tq_mavg_envelope_inverse replaces MovingAverageEnvelope.invert_representation
 * (representation.py:57-60): rep:[N,2*Cw,L] fp32 -> wave:[N,Cw,L] fp32.                                */
int tq_mavg_envelope_inverse(const float* rep, float* wave, int32_t N, int32_t Cw, int64_t L,
                             double log_eps, double eps, void* stream);


/* ---- forward representations (the step before the path: SURVEY 8(f) rank 2) ------------------------- *
 * tq_logspec_forward replaces LogSpectrogram.get_representation (tqdne/representation.py:140-150,
 * 163-169) with librosa 0.11 stft(x, n_fft, hop_length) semantics (center=True, zero padding, periodic
 * Hann): wave:[items, L] fp32 -> rep:[items, n_fft/2, frames] (Nyquist row dropped), frames = 1 + L/hop,
 * rep = 2*(log(max(|S|, clip)) - log(clip)) / (log_max - log(clip)) - 1.  precision: arithmetic
 * (TQ_F32 like librosa on float32 input, TQ_F64 like float64 input); rep_dtype: TQ_F32 or TQ_F64.      */
int tq_logspec_forward(const float* wave, void* rep, int32_t rep_dtype, int32_t items, int32_t n_fft,
                       int32_t hop, int64_t L, int32_t frames, double clip, double log_max,
                       int32_t precision, void* stream);
/* tq_mavg_envelope_forward replaces MovingAverageEnvelope.get_representation (representation.py:47-55):
 * wave:[N,Cw,L] fp32 -> rep:[N,2*Cw,L] fp32 = cat(x/(env+eps), log(env+log_eps) - log(log_eps)/2),
 * env = np.convolve(|x|, ones(window)/window, mode="same").                                              */
int tq_mavg_envelope_forward(const float* wave, float* rep, int32_t N, int32_t Cw, int64_t L,
                             int32_t window, double log_eps, double eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TQDNE_B200_H */
