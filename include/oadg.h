/*
 * oadg.h -- C ABI of libOADG.so: the B200 (sm_100a) implementation of the OA-DG
 * per-step hot path (OA-Mix transform + OA-Loss contrastive loss).
 *
 * The reference (WoojuLee24/OA-DG) is pure Python and has NO FFI of its own
 * (setup.py:220 ext_modules=[]); what it binds for this path are third-party
 * wheels.  Each entry point below therefore names the reference *call site*
 * whose native arithmetic it replaces.  INTEGRATION.md shows the ctypes stub a
 * maintainer would add on the reference side.
 *
 * Conventions
 *   - plain C: raw pointers + sizes, no torch / C++ types in any signature;
 *   - every pointer named *_dev is DEVICE memory owned by the caller, every
 *     pointer named *_host is host memory owned by the caller; the library never
 *     frees caller memory and keeps no global state besides lazily-set kernel
 *     attributes;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     all work is stream-ordered and asynchronous unless stated otherwise;
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = argument error (OADG_E_*);
 *     nothing throws;
 *   - re-entrant: may be called concurrently from several host threads on
 *     different streams with disjoint buffers.
 *   - images are u8, HWC, 3 channels, row pitch = 3*W bytes (tightly packed, the
 *     layout of the reference's numpy frames, loading.py:18-48).
 */
#ifndef OADG_H_
#define OADG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OADG_ABI_VERSION 1

#define OADG_E_ARG      (-1)   /* null pointer / bad size                         */
#define OADG_E_PLAN     (-2)   /* malformed plan blob; or a chain launch gave up waiting for work (views incomplete) */
#define OADG_E_LIMIT    (-3)   /* exceeds a compiled limit (OADG_MAX_*)           */
#define OADG_E_NOBOX    (-5)   /* OA-Mix sampler: no multi-level box could be placed
                                  (the reference raises ValueError from np.stack([]), oa_mix.py:217) */
#define OADG_E_ROWS     (-4)   /* OA-Loss: fewer than 2*ori_size rows (the reference
                                  raises RuntimeError at contrastive_loss.py:205) */

#define OADG_MAX_WIDTH    8    /* mixture_width  (reference default 3)            */
#define OADG_MAX_DEPTH    8    /* mixture_depth  (reference draws 1..3)           */
#define OADG_MAX_REGIONS  3    /* multi-level boxes (1..2) + the outside region   */

/* ---- op kinds: reference aug list, oa_mix.py:15-29 ------------------------ */
enum {
  OADG_OP_AUTOCONTRAST = 0,  /* augmix.py:64  PIL.ImageOps.autocontrast          */
  OADG_OP_EQUALIZE     = 1,  /* augmix.py:68  PIL.ImageOps.equalize              */
  OADG_OP_POSTERIZE    = 2,  /* augmix.py:72  ImageOps.posterize(bits = p0)      */
  OADG_OP_SOLARIZE     = 3,  /* augmix.py:103 ImageOps.solarize(thr = p0)        */
  OADG_OP_INVERT       = 4,  /* oa_mix.py:270-276 -warpAffine(+-1 px)  (tx=p0, ty=p1) */
  OADG_OP_COLOR        = 5,  /* augmix.py:192 ImageEnhance.Color(factor)         */
  OADG_OP_CONTRAST     = 6,  /* augmix.py:198 ImageEnhance.Contrast(factor)      */
  OADG_OP_BRIGHTNESS   = 7,  /* augmix.py:204 ImageEnhance.Brightness(factor)    */
  OADG_OP_SHARPNESS    = 8,  /* augmix.py:210 ImageEnhance.Sharpness(factor)     */
  OADG_OP_BG_AFFINE    = 9,  /* bbox_augmentation.py:240-302 bg_only_{rotate,shear,translate} */
  OADG_OP_BBO_AFFINE   = 10  /* bbox_augmentation.py:31-118 bboxes_only_{rotate,shear,translate} */
};

/* ---- plan records (host blob, copied verbatim to the device) --------------
 * All records are plain-old-data with natural alignment; Python packs them as
 * numpy structured arrays (oadg_b200/plan.py) and oadg_struct_sizes() lets the
 * binding verify the layout at load time.                                      */

typedef struct oadg_gt {        /* one gt box of one view: blurred-mask source (oa_mix.py:78-91) */
  int32_t lo[4];                /* x1,y1,x2,y2 on the (H/sr, W/sr) canvas, python-slice-resolved */
  int32_t blur;                 /* 0 => GaussianBlur skipped (sigma <= 0, oa_mix.py:89)          */
  int32_t kx, ky;               /* gaussian kernel sizes: cvRound(sigma*8+1)|1                   */
  int32_t view;                 /* owning view index                                             */
  double  sigma_x, sigma_y;
  int32_t supp[4];              /* conservative support rect of the mask [x0,y0,x1,y1) full res  */
} oadg_gt_t;

typedef struct oadg_op {        /* one OAMix.aug() draw (oa_mix.py:264-279)                      */
  int32_t kind;                 /* OADG_OP_*                                                     */
  int32_t p0, p1;               /* posterize bits / solarize threshold / invert tx,ty            */
  float   factor;               /* enhance factor                                                */
  double  minv[6];              /* BG_AFFINE: inverse (dst->src) 2x3, doubles, cv::warpAffine    */
  int32_t bbo_first, bbo_count; /* BBO_AFFINE: slice of the bbo record array                     */
  int32_t lut;                  /* slot in the LUT workspace (filled by the library), -1 if none */
  int32_t scratch;              /* BBO scratch image slot (filled by the library), -1 if none    */
} oadg_op_t;

typedef struct oadg_bbo {       /* one per (bboxes-only op, valid gt box), bbox_augmentation.py:45-71 */
  int32_t gt;                   /* global gt index (into the oadg_gt array)                      */
  int32_t pad;
  double  minv[6];              /* inverse affine of this box's random warp                      */
} oadg_bbo_t;

typedef struct oadg_target {    /* one object-aware mixing target (oa_mix.py:245-262,287-301)    */
  int32_t kind;                 /* 0 = blurred gt mask (gt = global gt index), 1 = hard box      */
  int32_t gt;
  int32_t box[4];               /* hard box x1,y1,x2,y2 (python-slice-resolved)                  */
  float   m_oa;                 /* np.float32(uniform(0,.5|1))  (oa_mix.py:296,298)              */
  int32_t pad;
} oadg_target_t;

typedef struct oadg_view {      /* one generated view (one oamix() call, oa_mix.py:207-243)      */
  int32_t H, W;
  int32_t img;                  /* index into the caller's source/output image pointer tables    */
  int32_t n_gt, gt_first;       /* slice of the oadg_gt array                                    */
  int32_t n_ml;                 /* multi-level boxes (regions = n_ml + 1, last = outside)        */
  int32_t ml_box[2][4];
  int32_t width;                /* mixture_width                                                 */
  int32_t depth[OADG_MAX_WIDTH];
  float   ws[OADG_MAX_WIDTH];   /* np.float32(dirichlet)  (oa_mix.py:212)                        */
  int32_t op_first;             /* ops of branch b, depth d, region r live at
                                   op_first + (b*OADG_MAX_DEPTH + d)*OADG_MAX_REGIONS + r         */
  int32_t n_tgt, tgt_first;     /* slice of the oadg_target array                                */
  int32_t pad;
  double  m;                    /* np.random.beta(1,1)  (oa_mix.py:282)                          */
} oadg_view_t;

typedef struct oadg_plan_header {
  int32_t magic;                /* 0x4F414447 "OADG"                                             */
  int32_t abi;                  /* OADG_ABI_VERSION                                              */
  int32_t n_views, n_gt, n_ops, n_bbo, n_tgt;
  int32_t max_h, max_w;         /* frame size every workspace image is dimensioned for           */
  int32_t off_views, off_gt, off_ops, off_bbo, off_tgt;   /* byte offsets from the blob start    */
  int32_t total_bytes;
  int32_t pad;
} oadg_plan_header_t;

/* ---- library ---------------------------------------------------------------- */
int  oadg_abi_version(void);
/* sizes of the six plan structs in the order header, view, gt, op, bbo, target */
void oadg_struct_sizes(int32_t out[6]);
const char* oadg_error_string(int code);
/* cudaMemcpyAsync between a host buffer (page-locked for a truly asynchronous copy) and a device buffer on `stream`;
 * to_device != 0: host -> device.  Host-side plumbing of the plugin (frame upload / view download), no reference
 * counterpart. */
int oadg_memcpy_async(void* dst, const void* src, size_t bytes, int to_device, void* stream);

/* ---- OA-Mix ----------------------------------------------------------------- */

/* ---- host-side plan sampler (no CUDA) ------------------------------------------------------------------
 * Draws everything random about n_img OAMix.oamix calls (oa_mix.py:207-262,281-298: Dirichlet weights,
 * multi-level boxes, depths, ops and their parameters, object-aware boxes, Beta / uniform mixing coefficients)
 * from the caller's random stream in the reference's draw order and packs the plan blob.  `rng` carries the
 * generator: for the reference it is the MT19937 bit generator behind numpy's global np.random (its
 * next_uint32 / next_double entry points), so np.random.seed(s) reproduces the reference's plan draw for draw.
 * scores[i]: the saliency score of every gt box of image i (oadg_saliency_scores; -1 for boxes narrower than 4).
 * Outputs besides the blob: the multi-level boxes ([n_img][2][4] int64, n_ml_out[i] valid) and object-aware
 * random boxes ([n_img][5][4] int64, n_oa_out[i] valid) the transform returns in its result dict, and
 * S(P) = sum of branch depths per image.  Returns OADG_E_NOBOX when an image got no multi-level box.          */
typedef struct oadg_rng {
  void*    state;
  uint32_t (*next_uint32)(void* state);
  double   (*next_double)(void* state);
} oadg_rng_t;

typedef struct oadg_sampler_cfg {   /* OAMix.__init__ keys (oa_mix.py:34-72) */
  int32_t version;                  /* 0 = 'augmix', 1 = 'augmix.all' (oa_mix.py:15-29) */
  int32_t severity, mixture_width, mixture_depth, spatial_ratio, score_thresh;
  double  random_box_scale[2], random_box_ratio[2], oa_random_box_scale[2], oa_random_box_ratio[2];
  double  sigma_ratio;
} oadg_sampler_cfg_t;

int oadg_oamix_sample_plan(const oadg_rng_t* rng, const oadg_sampler_cfg_t* cfg, int n_img,
                           const int32_t* hw /* [n_img][2] = H, W */,
                           const float* const* gt /* n_img x [n_gt][4] f32 */, const int32_t* n_gt,
                           const double* const* scores,
                           void* plan_out, size_t plan_cap, size_t* plan_bytes,
                           int64_t* ml_boxes_out, int32_t* n_ml_out,
                           int64_t* oa_boxes_out, int32_t* n_oa_out, int32_t* depth_sum_out);

/* Spectral-residual saliency score of every gt box.
 * Replaces oa_mix.py:107-110 (cv2.saliency.StaticSaliencySpectralResidual +
 * np.mean(uint8(map*255))).  boxes_dev: n x 5 int32 {img, x1, y1, x2, y2} with the
 * int32-truncated gt box (oa_mix.py:102); boxes with a side < 4 must not be passed
 * (the caller assigns them -1, oa_mix.py:103-105).  imgs_dev: table (in device
 * memory) of device pointers to u8 HWC frames; hw_dev: n_img x 2 int32 {H, W}. */
int oadg_saliency_scores(const uint8_t* const* imgs_dev, const int32_t* hw_dev,
                         const int32_t* boxes_dev, int n_boxes,
                         double* scores_dev, void* stream);

/* Bytes of device workspace oadg_oamix_execute needs for this plan. */
int oadg_oamix_workspace_bytes(const void* plan_host, size_t plan_bytes, size_t* out_bytes);

/* Execute a batch of OA-Mix views.
 * Replaces OAMix.oamix (oa_mix.py:207-243) minus its RNG draws, which stay on the
 * host and arrive as the plan: multi-level composite of random ops
 * (oa_mix.py:222-236; augmix.py / bbox_augmentation.py ops), branch mixing and
 * object-aware mixing (oa_mix.py:281-309).
 *   plan_host   : packed plan blob (oadg_plan_header_t + arrays), host memory
 *   src_dev     : table in HOST memory of n_img DEVICE pointers, u8 HWC sources
 *   dst_dev     : table in HOST memory of n_views DEVICE pointers, u8 HWC outputs
 *   workspace_dev / workspace_bytes : scratch from oadg_oamix_workspace_bytes
 *   launches_out: optional, receives the number of kernels launched            */
int oadg_oamix_execute(const void* plan_host, size_t plan_bytes,
                       const uint8_t* const* src_dev, int n_img,
                       uint8_t* const* dst_dev,
                       void* workspace_dev, size_t workspace_bytes,
                       int* launches_out, void* stream);

/* Opt-in epilogue of the mix kernel: besides the uint8 views, write what the rest of the training pipeline makes of
 * them -- Normalize (mmdet/datasets/pipelines/transforms.py:672-704 -> mmcv.imnormalize: BGR->RGB swap, float32
 * (x - mean) * (1/std) in cv2's arithmetic), Pad to a multiple of size_divisor with zeros (transforms.py:573-640) and
 * the HWC -> CHW transpose of DefaultFormatBundle (formating.py:217-234) -- as float32 [3, Hp, Wp] tensors, for the
 * generated views and (optionally) for the untouched source frames (the first view under keep_orig).  The values
 * are exact: a 3 x 256 table of the normalised value of every uint8 level is built on the host in float64 / float32
 * exactly as cv2.subtract / cv2.multiply round. */
typedef struct oadg_fused_out {
  float mean[3], std[3];        /* per OUTPUT channel (after the optional swap), as the Normalize config gives them */
  int32_t to_rgb;               /* != 0: channel k of the output is channel 2-k of the frame                        */
  int32_t size_divisor;         /* >= 1: Hp = ceil(H / d) * d, Wp likewise                                          */
  float* const* view_f32_dev;   /* HOST table of n_views DEVICE pointers, each [3, Hp, Wp] float32                  */
  float* const* src_f32_dev;    /* HOST table of n_img DEVICE pointers or NULL (entries may be NULL)                */
} oadg_fused_out_t;

int oadg_oamix_execute_fused(const void* plan_host, size_t plan_bytes,
                             const uint8_t* const* src_dev, int n_img,
                             uint8_t* const* dst_dev, const oadg_fused_out_t* fused,
                             void* workspace_dev, size_t workspace_bytes,
                             int* launches_out, void* stream);

/* Same as oadg_oamix_execute with the persistent chain kernel limited to ctas_per_sm resident CTAs per SM
 * (0 = as many as fit, i.e. oadg_oamix_execute).  Two such launches on different streams, each with its own
 * workspace, share the SMs: the tiles of one batch fill the dependency stalls of the other (the reference has no
 * counterpart; its batches are independent DataLoader items, oa_mix.py:187-204). */
int oadg_oamix_execute_shared(const void* plan_host, size_t plan_bytes,
                              const uint8_t* const* src_dev, int n_img,
                              uint8_t* const* dst_dev,
                              void* workspace_dev, size_t workspace_bytes,
                              int ctas_per_sm, int* launches_out, void* stream);

/* Same as oadg_oamix_execute, with CUDA events on `stream` around the two launches (the chain kernel and the mix
 * kernel); the call synchronises the stream before returning (measurement only).  n_items_out / n_tiles_out: size
 * of the chain kernel's work queue.  kind_stats (optional, 48 x uint64): CTA-busy nanoseconds [16], tiles [16] and
 * longest tile ns [16] per work-item kind (0 profile, 1 mask, 2 hist, 3 lut, 4 frame copy, 5 bbo blend, 6 bbo
 * catch-up, 7 / 8 / 9 depth-step tiles: streaming / staged bg-only / mixed; busy slot 10 = ns spent waiting for
 * dependencies). */
int oadg_oamix_execute_profiled(const void* plan_host, size_t plan_bytes,
                                const uint8_t* const* src_dev, int n_img,
                                uint8_t* const* dst_dev,
                                void* workspace_dev, size_t workspace_bytes,
                                float* ms_chain, float* ms_mix, int* n_items_out, int* n_tiles_out,
                                unsigned long long* kind_stats, void* stream);

/* A chain-kernel CTA that finds no ready work for 2 s (inconsistent dependency tables, a stalled device) leaves
 * the kernel and raises a sticky device flag instead of hanging the GPU; the flag of every launch is copied to
 * page-locked memory behind the launch.  This call reports it for the launches of this process so far:
 * wait != 0 waits for all of them, wait == 0 looks at the finished ones.  Returns OADG_E_PLAN when a launch left
 * views incomplete (the same code the next oadg_oamix_execute* call would return), else 0. */
int oadg_oamix_poll_fault(int wait);

/* Measurement aid (set OADG_TRACE=1): per work item of the last profiled execution {kind, obj, tiles, dependency
 * count, first claim / last publish in us, first 8 dependencies}; returns the number of items. */
int oadg_oamix_last_trace(void* out, int cap);

/* ---- OA-Loss ---------------------------------------------------------------- */

/* Workspace bytes for forward+backward at N rows, C channels. */
int oadg_supcon_workspace_bytes(int n, int c, size_t* out_bytes);

/* Forward of ContrastiveLossPlus (contrastive_loss_plus.py:31-50 ->
 * contrastive_loss.py:170-232 supcontrast -> :147-167 supcontrast_mask):
 * double L2 normalisation, similarity / temperature, masked InfoNCE.
 *   feats_dev   : [n, c] f32 row-major (pre-normalisation RoI embeddings)
 *   labels_dev  : [n] int64, already padded for random-proposal rows
 *   pair_dev    : [n] int32 other-view row of each row or -1
 *                 (the reference's hard-wired layout, or a generalised map)
 *   loss_dev    : [1] f32 out (loss_weight applied); 0 when #fg <= min_samples
 *   workspace   : keeps the normalised embeddings and per-row statistics for bwd */
int oadg_supcon_forward(const float* feats_dev, const int64_t* labels_dev,
                        const int32_t* pair_dev, int n, int c,
                        float temperature, float loss_weight, int min_samples,
                        int normalized_input,
                        float* loss_dev, void* workspace_dev, size_t workspace_bytes,
                        int* launches_out, void* stream);

/* Backward: grad_feats_dev[n,c] = dloss/dfeats * (*grad_loss_dev).  Must follow a
 * forward on the same workspace. */
int oadg_supcon_backward(const float* feats_dev, const int64_t* labels_dev,
                         const int32_t* pair_dev, int n, int c,
                         float temperature, float loss_weight, int normalized_input,
                         const float* grad_loss_dev, float* grad_feats_dev,
                         void* workspace_dev, size_t workspace_bytes,
                         int* launches_out, void* stream);

/* ---- OA-Loss across ranks (new capability; the reference's loss is per-rank local) --------------------
 * Anchors are this rank's rows [row0, row0 + n_rows) of the all-gathered, doubly-normalised embeddings
 * fhat_all [n_total, c]; contrasts are all n_total rows.  The caller (oadg_b200/distributed.py) performs the
 * collectives with torch.distributed / NCCL between the calls:
 *   normalize (local) -> all_gather(fhat, labels) -> forward_gathered -> all_reduce(loss), all_gather(stats)
 *   -> backward_gathered.  One workspace sized by oadg_supcon_workspace_bytes(n_total, c) spans the sequence. */
int oadg_supcon_normalize(const float* feats_dev, int n_rows, int n_total, int c, int normalized_input,
                          float* fhat_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
int oadg_supcon_forward_gathered(const float* fhat_all_dev, const int64_t* labels_all_dev,
                                 const int32_t* pair_all_dev, int n_total, int row0, int n_rows, int c,
                                 float temperature, float loss_weight, int min_samples,
                                 float* loss_part_dev, float* stats_local_dev /* [n_rows, 4] */,
                                 void* workspace_dev, size_t workspace_bytes, int* launches_out, void* stream);
int oadg_supcon_backward_gathered(const float* feats_local_dev, const int64_t* labels_all_dev,
                                  const int32_t* pair_all_dev, const float* stats_all_dev /* [n_total, 4] */,
                                  int n_total, int row0, int n_rows, int c, float temperature,
                                  int normalized_input, const float* grad_loss_dev, float* grad_feats_dev,
                                  void* workspace_dev, size_t workspace_bytes, int* launches_out, void* stream);

/* The same sequence with PACKED buffers, so that no pack / slice / copy kernels sit around the two collectives:
 *   gather_pack      normalises the rank's rows straight into its send buffer: n_rows rows of oadg_supcon_pack_width(c)
 *                    floats = [fhat | the row's int64 label as two 32-bit words | padding]; labels_dev may hold fewer
 *                    than n_rows labels, the rest take the last one (contrastive_loss_plus.py:44-47)
 *   -> all_gather(send) -> forward_packed reads the gathered buffer in place and writes the rank's TAIL buffer:
 *                    n_rows rows of 4 floats (row statistics) + one row {loss part, -, -, -}
 *   -> all_gather(tail) -> finish_packed: total loss (rank order, same bits on every rank) and the statistics of all
 *                    rows, kept in the workspace
 *   -> backward_packed (labels and statistics come from the workspace). */
int oadg_supcon_pack_width(int c);
int oadg_supcon_gather_pack(const float* feats_dev, const int64_t* labels_dev, int n_labels, int n_rows, int n_total,
                            int c, int normalized_input, float* send_dev, void* workspace_dev, size_t workspace_bytes,
                            void* stream);
int oadg_supcon_forward_packed(const float* recv_dev, const int32_t* pair_all_dev, int n_total, int row0, int n_rows,
                               int c, float temperature, float loss_weight, int min_samples, float* tail_dev,
                               void* workspace_dev, size_t workspace_bytes, int* launches_out, void* stream);
int oadg_supcon_finish_packed(const float* tail_all_dev, int world, int n_rows, int c, float* loss_dev,
                              void* workspace_dev, size_t workspace_bytes, void* stream);
int oadg_supcon_backward_packed(const float* feats_local_dev, const int32_t* pair_all_dev, int n_total, int row0,
                                int n_rows, int c, float temperature, int normalized_input, const float* grad_loss_dev,
                                float* grad_feats_dev, void* workspace_dev, size_t workspace_bytes, int* launches_out,
                                void* stream);

/* ---- OA-Loss, consistency half ------------------------------------------------------------------------------
 * Jensen-Shannon divergence between the two views' class distributions (cross_entropy_loss_plus.py:264-319
 * jsdv1_3_2aug): pred_dev [2 n, c] float32 = view 1 rows then view 2 rows, c <= 32 (c == 1: the RPN's single logit,
 * classes (sigmoid, 1 - sigmoid); else softmax).  Writes loss_dev[0] = the sum of the row divergences (the
 * reference's `/ len(p_aug1)` divides by the leading 1 of a [1, n, C] reshape) and
 * grad_dev [2 n, c] = d loss / d pred (so the backward is a scale).  scratch_dev: oadg_jsd2_scratch_bytes() bytes,
 * zeroed once by the caller. */
int oadg_jsd2_scratch_bytes(void);
int oadg_jsd2_forward(const float* pred_dev, int n, int c, float* loss_dev, float* grad_dev, void* scratch_dev,
                      void* stream);

/* ---- OA-Loss across ranks without a collective library on the critical path (new capability) ----------
 * Every rank owns one exportable device buffer (oadg_peer_alloc), maps the other ranks' buffers through CUDA IPC
 * (oadg_peer_export / _import, handles exchanged once by the host) and from then on the kernels of a step store their
 * results straight into every rank's buffer over NVLink and raise a per-source flag word there; the consumer waits on
 * its own flags (oadg_peer_wait).  No kernel of another library has to be resident on either side.
 *   oadg_supcon_gather_pack_peers = oadg_supcon_gather_pack whose packed rows land in EVERY rank's gather buffer
 *                                   (peers->base[r] + rows_offset, row index rank * n_rows), followed by
 *                                   flag word (peers->base[r] + flag_offset)[rank] = seq
 *   oadg_peer_scatter              copies `bytes` (multiple of 16) from this rank's own buffer at `offset` to the same
 *                                   offset of every other rank, then raises flag word [rank] at flag_offset
 *   oadg_peer_wait                 returns (in stream order) once flags_dev[r] has reached seq for every r < world
 *                                   (cyclic comparison); after timeout_ms it writes 1 to *fault_host (page-locked, mapped)
 *                                   and gives up, so a dead peer surfaces as an error instead of a hang.
 * Next to its flag every sender leaves a tag (flag word [OADG_PEER_MAX + rank]; the pack kernel sends its row count,
 * oadg_peer_scatter whatever the caller passes): a waiter whose own tag differs writes 2 to *fault_host -- ranks that
 * disagree on the shape of the exchange must not read each other's rows at the wrong offsets.
 * counter_offset names a zero-initialised 32-bit word of the caller's own buffer (last-block-done ticket). */
#define OADG_PEER_MAX 16
typedef struct {
  void* base[OADG_PEER_MAX]; /* base[rank] is this rank's own buffer */
  int32_t world, rank;
} oadg_peers_t;
int oadg_peer_alloc(size_t bytes, void** ptr_out);
int oadg_peer_free(void* ptr);
int oadg_peer_export(void* ptr, unsigned char handle_out[64]);
int oadg_peer_import(const unsigned char handle[64], void** ptr_out);
int oadg_peer_release(void* ptr);
int oadg_peer_fault_alloc(uint32_t** fault_host_out);
int oadg_supcon_gather_pack_peers(const float* feats_dev, const int64_t* labels_dev, int n_labels, int n_rows,
                                  int n_total, int c, int normalized_input, const oadg_peers_t* peers,
                                  size_t rows_offset, size_t flag_offset, size_t counter_offset, uint32_t seq,
                                  void* workspace_dev, size_t workspace_bytes, void* stream);
int oadg_peer_scatter(const oadg_peers_t* peers, size_t offset, size_t bytes, size_t flag_offset,
                      size_t counter_offset, uint32_t seq, uint32_t tag, void* stream);
int oadg_peer_wait(const uint32_t* flags_dev, int world, uint32_t seq, uint32_t tag, uint32_t timeout_ms,
                   uint32_t* fault_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OADG_H_ */
