/*
 * orca_b200.h -- C ABI of liborca_b200.so: the B200-native (sm_100a) forward path of
 * Orca's multiscale genome-interaction predictor.
 *
 * Nothing equivalent exists in the reference (it is pure Python on torch.nn); each entry
 * point below replaces one `nn.Module.forward` of /root/reference/orca_modules.py and is
 * what a reference-side binding (ctypes) calls:
 *
 *   orca_b200_encoder_forward     <- Encoder.forward      orca_modules.py:929-980 (layers :811-927)
 *   orca_b200_encoder_forward_packed  the same + the one-hot feeder, selene_utils2.py:125-128 / :216-230
 *   orca_b200_encoder2_forward    <- Encoder2.forward     orca_modules.py:1151-1169
 *                                    Encoder2b.forward    orca_modules.py:1266-1276
 *                                    Encoder3.forward     orca_modules.py:1388-1406
 *   orca_b200_decoder_forward     <- Decoder.forward      orca_modules.py:461-488
 *                                    Decoder_1m.forward   orca_modules.py:782-800
 *   orca_b200_net_forward         <- Net.forward          orca_modules.py:1833-1900
 *   orca_b200_background_forward  <- block-nanmean + log of the caller's background matrix,
 *                                    orca_predict.py:693-697 and :724-737
 *   orca_b200_background_assemble <- the background matrix of a multi-region input,
 *                                    orca_predict._retrieve_multi, orca_predict.py:936-965
 *
 * Conventions
 *   - plain C types only; every tensor is a raw DEVICE pointer to fp32 unless stated
 *     otherwise; shapes and strides (in ELEMENTS) are explicit.
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation.  Kernel selection and precision are properties of a module HANDLE
 *     (orca_b200_module_set_option), never of the process: forwards on different handles are
 *     independent and reentrant across host threads and streams.  The only process-wide state is
 *     the launch counter (atomic), the thread-local last-error string and the opt-in, single-threaded
 *     benchmarking recorder (orca_b200_profile_enable).
 *   - the caller owns inputs, outputs and workspace (allocated through its own
 *     allocator, e.g. PyTorch's); the library owns only the packed weights inside a
 *     module handle (released by orca_b200_module_destroy).
 *   - return value: 0 on success, a negative ORCA_B200_E* code otherwise;
 *     orca_b200_last_error() describes the last failure on the calling thread.
 *   - there is no CPU fallback: host pointers passed as tensors are an error.
 */
#ifndef ORCA_B200_H
#define ORCA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORCA_B200_OK 0
#define ORCA_B200_EINVAL (-1)       /* bad argument / shape / architecture mismatch */
#define ORCA_B200_EUNSUPPORTED (-2) /* valid but not implemented for this configuration */
#define ORCA_B200_ECUDA (-3)        /* a CUDA runtime call failed */
#define ORCA_B200_EWORKSPACE (-4)   /* workspace smaller than *_workspace_bytes */

/* Module kinds: one per network class of orca_modules.py. */
enum {
  ORCA_B200_ENCODER = 1,    /* Encoder    orca_modules.py:803-980   (28 convs) */
  ORCA_B200_ENCODER2 = 2,   /* Encoder2   orca_modules.py:984-1169  (40 convs) */
  ORCA_B200_ENCODER2B = 3,  /* Encoder2b  orca_modules.py:1173-1276 (20 convs) */
  ORCA_B200_ENCODER3 = 4,   /* Encoder3   orca_modules.py:1279-1406 (24 convs) */
  ORCA_B200_DECODER = 5,    /* Decoder    orca_modules.py:16-488    (122 convs, 118 used per call) */
  ORCA_B200_DECODER_1M = 6, /* Decoder_1m orca_modules.py:491-800   (78 convs) */
  ORCA_B200_NET = 7         /* Net        orca_modules.py:1409-1900 (106 convs [+2 final_1d]) */
};
/*
 * Decoder / Decoder_1m / Net also cover the multi-map variants of orca_leukemia.py (Decoder(num_2d) :512-993,
 * Decoder_1m(num_2d) :996-1315, Net(num_2d, num_1d) :16-509: the same trees with `final` 64 -> max(num_2d,5) ->
 * num_2d and combiner inputs of 64+num_2d / 128+num_2d channels).  num_2d (1..8) is read off the module's own
 * final 1x1 convolution at creation; every tensor documented below as (B, 1, ...) is then (B, num_2d, ...).
 */

/* flags for orca_b200_module_create */
#define ORCA_B200_UPSAMPLE_NEAREST 0u  /* Decoder(upsample_mode='nearest')  */
#define ORCA_B200_UPSAMPLE_BILINEAR 1u /* Decoder(upsample_mode='bilinear'), align_corners=False */

/* compute path selection (ORCA_B200_OPT_IMPL); both are CUDA, there is no CPU path */
#define ORCA_B200_IMPL_AUTO 0  /* tcgen05 tensor-core kernels where available, else SIMT */
#define ORCA_B200_IMPL_SIMT 1  /* fp32 CUDA-core implicit-GEMM everywhere */
#define ORCA_B200_IMPL_TC 2    /* require the tcgen05 path (error if a layer lacks it) */

/*
 * One convolution of a module, in the reference's own parameter layout (HOST pointers,
 * fp32, exactly the tensors of the reference state_dict):
 *   weight  (c_out, c_in, kh, kw) row-major  [Conv1d: kh = 1, kw = 9; Conv2d: 3x3 or 1x1]
 *   bias    (c_out)
 *   bn_*    (c_out) of the BatchNorm that directly follows the conv in its nn.Sequential,
 *           all four NULL when no BatchNorm follows (e.g. Encoder2.downblocks[*][3],
 *           orca_modules.py:1115-1120).  Eval-mode semantics (running stats).
 */
typedef struct orca_b200_conv_params {
  int32_t c_in, c_out, kh, kw, dilation;
  const float* weight;
  const float* bias;
  const float* bn_weight;
  const float* bn_bias;
  const float* bn_mean;
  const float* bn_var;
  float bn_eps;
} orca_b200_conv_params;

typedef struct orca_b200_module orca_b200_module;

/* library / device */
const char* orca_b200_version(void);
const char* orca_b200_last_error(void);
/*
 * Per-handle options.
 *   ORCA_B200_OPT_IMPL                 ORCA_B200_IMPL_* (default AUTO).
 *   ORCA_B200_OPT_ENCODER_FP16_STAGES  Encoder / Net handles: the first n of the Encoder's 7 stages
 *       (orca_modules.py:811-927) run their convolutions as ONE fp16 tensor-core product with fp16 activations; the
 *       remaining stages -- and every other module -- keep fp32-grade arithmetic (operands split into two bf16, three
 *       products).  Default 4: stages 1-4 hold 98.7 % of the encoder FLOP and their 2^-12 rounding noise is averaged out
 *       by the pooling and convolutions of stages 5-7 -- PROVIDED the folded weights are well conditioned: module_create measures
 *       the spread of the folded weight row norms of those stages (max / median per conv) and the default drops to 0 when
 *       it exceeds 4 (BatchNorm scales spanning orders of magnitude).  0 = three products everywhere; -1 restores the
 *       default; orca_b200_module_get_option returns the effective value.
 * orca_b200_module_status reads (and optionally clears) the handle's device status word; it synchronises the device.
 *   bit 0  ORCA_B200_STATUS_FP16_RANGE: a value written by a single-pass fp16 stage exceeded the fp16 range guard
 *          (|x| > 60000) since the word was last cleared -- the output of that forward is not trustworthy; rerun it
 *          with ORCA_B200_OPT_ENCODER_FP16_STAGES = 0 (orca_b200.modules does this automatically).
 */
#define ORCA_B200_OPT_IMPL 1
#define ORCA_B200_OPT_ENCODER_FP16_STAGES 2
#define ORCA_B200_STATUS_FP16_RANGE 1u
int orca_b200_module_set_option(orca_b200_module* m, int option, int value);
int orca_b200_module_get_option(const orca_b200_module* m, int option);
int orca_b200_module_status(const orca_b200_module* m, uint32_t* status, int32_t clear);

/* number of kernel launches issued by this library since load (all threads) */
uint64_t orca_b200_launch_count(void);

/*
 * Per-launch timing of the convolution kernels (bench.py's roofline leg).  While enabled every
 * conv launch is bracketed by CUDA events on its own stream; orca_b200_profile_summary
 * synchronises on them and writes a JSON array of {c_in, c_out, taps, dil, tc, launches, ms,
 * flop} aggregates into buf (returns the number of bytes needed, or a negative error).
 * Enabling/disabling clears the records.  Not thread-safe; meant for one benchmarking thread.
 */
int orca_b200_profile_enable(int on);
int64_t orca_b200_profile_summary(char* buf, int64_t cap);

/*
 * Build a module: folds BatchNorm into each conv, packs the weights into the kernel
 * layouts and uploads them to the CURRENT CUDA device.  `convs` lists the module's
 * convolutions in the order of the reference class definition (= state_dict order);
 * the library checks every (c_in, c_out, kh, kw, dilation) against the architecture of
 * `kind` and fails with ORCA_B200_EINVAL on any mismatch.
 * `num_1d` (Net only): channels of the final_1d head, 0 if absent.
 */
int orca_b200_module_create(int kind, const orca_b200_conv_params* convs, int32_t n_convs,
                            uint32_t flags, int32_t num_1d, orca_b200_module** out);
void orca_b200_module_destroy(orca_b200_module* m);
int orca_b200_module_kind(const orca_b200_module* m);
int orca_b200_module_num_2d(const orca_b200_module* m); /* output maps per sample (1 for orca_modules) */

/*
 * Encoder.forward.  x: (B, 4, L) with element strides (sB, sC, sL) -- genomepredict hands
 * over a transposed view with strides (4L, 1, 4) (orca_predict.py:333-337).  L must be a
 * positive multiple of 4000.  Strides may be negative: the reverse-complement strand
 * RC(x)[b,c,l] = x[b,3-c,L-1-l] (orca_predict.py:324-329) is x + 3*sC + (L-1)*sL walked with
 * (sB, -sC, -sL).  out: (B, L/4000, 128) channel-last, contiguous.
 * Internally processed in chunks of `chunk_bp` (+112,000 bp halo per interior side,
 * the reference's own overlap, orca_modules.py:931-932); chunk_bp = 0 picks the default.
 * [bin_begin, bin_end) selects the 4 kb bins to produce (sequence sharding across
 * GPUs); pass 0, L/4000 for all.  `out` always addresses bin 0 of sample 0.
 * A shard may hold only a window of the sequence: x then addresses position `x_pos0` and
 * covers `x_len` positions (pass 0, L for the whole sequence); the window must contain
 * [bin_begin*4000 - 112008, bin_end*4000 + 112012) clipped to [0, L).
 */
size_t orca_b200_encoder_workspace_bytes(const orca_b200_module* m, int64_t B, int64_t L,
                                         int64_t chunk_bp);
int orca_b200_encoder_forward(const orca_b200_module* m, const float* x, int64_t B, int64_t L,
                              int64_t sB, int64_t sC, int64_t sL, int64_t x_pos0, int64_t x_len,
                              float* out, int64_t bin_begin, int64_t bin_end, int64_t chunk_bp,
                              void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same forward from PACKED bases, one byte per position -- replaces the reference's
 * host-side one-hot feeder + fp32 upload (selene_utils2.py:125-128, :216-230; orca_predict.py:324-337:
 * 16 B/bp over PCIe per strand and model) by 1 B/bp uploaded once.
 *   bases: (B, L) bytes with BYTE strides (sB, sL).  A byte is either a code 0..4 = A, C, G, T, N or raw
 *          ASCII as in a FASTA record ('A','C','G','T' in either case; anything else is N).  A/C/G/T become the
 *          one-hot rows of the reference's ACGT channel order, N becomes 0.25 in all four channels.
 *   complement != 0 with sL < 0 reads the reverse-complement strand in place: pass bases + (L-1)*|sL|
 *          (the last position) and a negative sL; codes are complemented (A<->T, C<->G) on the fly.
 * Everything else as for orca_b200_encoder_forward (x_pos0 / x_len are in positions of the walked strand).
 */
int orca_b200_encoder_forward_packed(const orca_b200_module* m, const uint8_t* bases, int64_t B,
                                     int64_t L, int64_t sB, int64_t sL, int32_t complement,
                                     int64_t x_pos0, int64_t x_len, float* out, int64_t bin_begin,
                                     int64_t bin_end, int64_t chunk_bp, void* workspace,
                                     size_t workspace_bytes, void* stream);

/*
 * Encoder2 / Encoder2b / Encoder3 forward.  x: (B, 128, P) with element strides
 * (sB, sC, sL); P divisible by 32 (Encoder2/2b) or 8 (Encoder3).
 * outs: n_out device pointers, finest -> coarsest, each (B, P >> i, 128) channel-last:
 *   Encoder2: 6 (P .. P/32)   Encoder2b: 6 (outs[0] is a copy of x)   Encoder3: 4 (P .. P/8)
 * `coarsest_only` != 0 (Encoder2 only): compute just the pooling half and write only
 * outs[n_out-1] -- what genomepredict_256Mb consumes (`net1(...)[-1]`, orca_predict.py:675-683).
 */
size_t orca_b200_encoder2_workspace_bytes(const orca_b200_module* m, int64_t B, int64_t P);
int orca_b200_encoder2_forward(const orca_b200_module* m, const float* x, int64_t B, int64_t P,
                               int64_t sB, int64_t sC, int64_t sL, float* const* outs,
                               int32_t n_out, int32_t coarsest_only, void* workspace,
                               size_t workspace_bytes, void* stream);

/*
 * Decoder.forward(x, distenc, y) and Decoder_1m.forward(x).          C = num_2d of the module (1 in orca_modules)
 *   x       (B, 128, S), element strides (xsB, xsC, xsL)           S = 250 in Orca
 *   distenc (B, C, S, S), strides (dsB, dsC, dsH, dsW) [dsB may be 0: expanded view]; Decoder only
 *   y       (B, C, S/2, S/2), strides (ysB, ysC, ysH, ysW) or NULL; Decoder only
 *   out     (B, C, S, S) contiguous
 */
size_t orca_b200_decoder_workspace_bytes(const orca_b200_module* m, int64_t B, int64_t S);
int orca_b200_decoder_forward(const orca_b200_module* m, const float* x, int64_t B, int64_t S,
                              int64_t xsB, int64_t xsC, int64_t xsL, const float* distenc,
                              int64_t dsB, int64_t dsC, int64_t dsH, int64_t dsW, const float* y,
                              int64_t ysB, int64_t ysC, int64_t ysH, int64_t ysW, float* out,
                              void* workspace, size_t workspace_bytes, void* stream);

/*
 * Net.forward (Orca-1Mb).  x as for the encoder; L/4000 is the map size S.
 * out: (B, 1, S, S); out_1d: (B, num_1d, S) contiguous or NULL when the module has no
 * final_1d head (or the caller discards it, as H1esc_1M.forward does, orca_models.py:491-494).
 */
size_t orca_b200_net_workspace_bytes(const orca_b200_module* m, int64_t B, int64_t L);
int orca_b200_net_forward(const orca_b200_module* m, const float* x, int64_t B, int64_t L,
                          int64_t sB, int64_t sC, int64_t sL, float* out, float* out_1d,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Net.forward from packed bases (see orca_b200_encoder_forward_packed). */
int orca_b200_net_forward_packed(const orca_b200_module* m, const uint8_t* bases, int64_t B, int64_t L,
                                 int64_t sB, int64_t sL, int32_t complement, float* out,
                                 float* out_1d, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Background distance-matrix level (orca_predict.py:724-737 then :693-697):
 *   out[i][j] = log( nanmean( normmat[r0 + i*f : r0 + (i+1)*f, r0 + j*f : r0 + (j+1)*f] ) )
 * normmat: (n, n) fp64 row-major on the device; out: (S, S) fp32; requires r0 + S*f <= n.
 * `flip` != 0 writes out[S-1-i][S-1-j] instead (reverse strand, orca_predict.py:703).
 * nanmean is evaluated as the reference does: mean over the inner axis first, then
 * over the outer axis, NaN entries skipped at each step.
 */
int orca_b200_background_forward(const double* normmat, int64_t n, int64_t r0, int64_t f,
                                 int64_t S, int32_t flip, float* out, void* stream);
/* Same, and additionally (out_mean != NULL) the float64 block nanmean itself, (S, S) row-major and NOT flipped: the
 * matrices genomepredict_256Mb returns as output['normmats'] (orca_predict.py:724-737). */
int orca_b200_background_level(const double* normmat, int64_t n, int64_t r0, int64_t f, int64_t S,
                               int32_t flip, float* out_log, double* out_mean, void* stream);

/*
 * Background matrix of a multi-region 256 Mb input, assembled on the device -- replaces the normmat branch of
 * orca_predict._retrieve_multi (orca_predict.py:936-965), which builds the (8000, 8000) float64 matrix with numpy
 * on the host (and the caller then uploads 512 MB):
 *   block(a, b) = cis[(|acoor[:, None] - bcoor[None, :]| / binsize).astype(int)]   same chromosome,
 *                 acoor = np.linspace(start, end, nb + 1)[:-1],  nb = int((end - start) / binsize)
 *               = trans                                                          different chromosomes
 *   rows of a reverse-strand region a and columns of a reverse-strand region b are reversed.
 * regions: HOST array; chrom is any integer id (equal ids = same chromosome).  cis: DEVICE float64 curve of
 * n_cis entries (may hold NaN pads, orca_models.py:626-633).  out: DEVICE (n, n) float64 row-major with
 * n = orca_b200_background_bins(...).  workspace: DEVICE scratch of at least 12*n + 512 bytes.
 * The call synchronises `stream` once (a 12*n-byte table upload) before the fill kernel is enqueued.
 */
typedef struct orca_b200_region {
  int32_t chrom;
  int32_t reverse; /* 1 for a '-' strand region */
  int64_t start, end;
} orca_b200_region;
int64_t orca_b200_background_bins(const orca_b200_region* regions, int32_t n_regions, int64_t binsize);
int orca_b200_background_assemble(const orca_b200_region* regions, int32_t n_regions, const double* cis,
                                  int64_t n_cis, double trans, int64_t binsize, double* out, int64_t n,
                                  void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ORCA_B200_H */
