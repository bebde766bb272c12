/* piccolo_b200 — C ABI of the B200-native PICCOLO sampling-loss pose search.
 *
 * Drop-in boundary.  The reference (82magnolia/piccolo) is pure Python and has no FFI; its hot
 * path is the torch op chain behind these call sites, which the entry points below replace:
 *
 *   pcl_cloud_create   the per-room upload `xyz.to(device)`, `rgb.to(device)`  (localize.py:159-164)
 *                      + the clamp box `quantile()`                             (utils.py:208-229,
 *                        omniloc.py:53-55, :245-247)
 *   pcl_image_create   the per-query upload `img.to(device)`                    (localize.py:169-170, :212-213)
 *   pcl_score          the T×R forward-only loop of `trim_input_loss`           (utils.py:484-499)
 *                      and `sampling_loss`                                      (omniloc.py:105-157)
 *   pcl_score_grid     the same loop, given as translations x rotations        (utils.py:484-499)
 *   pcl_topk           `loss_table.flatten().argsort()[:num_input]`             (utils.py:501-502)
 *   pcl_loss_fwd_bwd   `SamplingLoss.forward` + autograd backward               (omniloc.py:171-202)
 *                      `BatchSamplingLoss.forward` + backward                   (omniloc.py:311-356)
 *   pcl_color_*        `color_match`, `color_mod` (per-query colour preprocessing)      (color_utils.py:146-234, :7-65)
 *   pcl_refine_*       the optimisation loops of `omniloc` / `omniloc_batch`:
 *                      loss, backward, Adam.step, ReduceLROnPlateau.step, clamp (omniloc.py:44-58, :249-269)
 *
 * Conventions
 *   - every pointer named *_dev is a CUDA device pointer on the current device; everything else is
 *     host memory.  All tensors are float32, contiguous, row-major.
 *   - a pose is 6 floats (tx, ty, tz, yaw, pitch, roll); R = Rz(yaw)·Ry(pitch)·Rx(roll)
 *     (utils.py:425-453), camera-frame point q = R (p - t).
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls are asynchronous with
 *     respect to the host unless stated otherwise; results are ordered on `stream`.
 *   - every function returns 0 on success or a negative pcl_status; pcl_last_error() returns a
 *     thread-local message for the last failure.  No C++ exception crosses the ABI.
 *   - handles (pcl_cloud, pcl_image) are immutable after creation and may be shared by streams;
 *     a pcl_refine is single-owner.  The library never frees caller memory.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     PCL_ERR_CUDA.
 */
#ifndef PICCOLO_B200_H
#define PICCOLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCL_ABI_VERSION 2

typedef enum pcl_status {
  PCL_OK = 0,
  PCL_ERR_INVALID = -1,   /* bad argument */
  PCL_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
  PCL_ERR_FORMAT = -3     /* image cannot be represented in the requested texel format */
} pcl_status;

/* texel formats of a pcl_image */
#define PCL_IMAGE_AUTO 0   /* image exactly k/255: F16D up to 2048x4096 (+ a compact U8Q/U8P companion for small gradient batches), U8Q beyond; otherwise F32 */
#define PCL_IMAGE_U8Q 1    /* 16-byte footprint entries {nw,ne,sw,se} RGBA8: one 128-bit load per sample */
#define PCL_IMAGE_F32 2    /* fp32 RGBA texels: arbitrary float images */
#define PCL_IMAGE_U8P 3    /* plain RGBA8 texels (4 B/texel): smallest table, four 32-bit loads */
#define PCL_IMAGE_TEX 4    /* RGBA8 cudaArray behind a texture object: three 2x2 gathers per sample */
#define PCL_IMAGE_F16D 5   /* 32-byte footprint entries of fp16 bilinear-basis values: two 128-bit loads, no unpack arithmetic */

/* point ordering of a pcl_cloud */
#define PCL_CLOUD_KEEP_ORDER 0
#define PCL_CLOUD_MORTON 1   /* sort by 3-D Morton code: neighbouring lanes gather neighbouring texels */

typedef struct pcl_cloud pcl_cloud;
typedef struct pcl_image pcl_image;
typedef struct pcl_refine pcl_refine;
typedef struct pcl_comm pcl_comm;

int pcl_abi_version(void);
const char* pcl_last_error(void);
/* number of kernels launched by this library in the calling process (all threads) since load */
int64_t pcl_launch_count(void);

/* Tuning knobs for A/B measurements and tests (process-global; defaults are read once from PCL_<NAME> in the
 * environment; value < 0 restores the default): PERSIST (1: small refinement batches run all iterations in one
 * cooperative launch), PDL (programmatic dependent launch of per-iteration launches), PB_FWD / PB_BWD (poses per
 * CTA of the generic kernels, 0 = auto), WAVES, SWAP, GRID_SWAP (block order), SMALL_TABLE (compact texel table for
 * small gradient batches), RF_NPB (candidates per pose block of the fused refinement, 0 = auto),
 * RF_DEBUG (1: record per-warp cycle counters, see pcl_refine_debug_stats). */
int pcl_set_option(const char* name, int value);

/* ---- coloured point cloud ------------------------------------------------------------------ */
/* xyz_n3_dev, rgb_n3_dev: (N,3) float32.  q: out_of_room_quantile (clamp box = order statistics
 * int(N*q) and int(N*(1-q)) per axis, utils.py:222-227).  Asynchronous: all work is ordered on `stream`, and the
 * storage is freed in stream order there.  Compute entries may use the handle on another stream once that stream is
 * ordered after the creation (event / stream wait): every entry records its use, and pcl_cloud_destroy /
 * pcl_image_destroy make the creation stream wait for the most recent record before the storage returns to the pool
 * (uses on several foreign streams: order the earlier ones before the last one yourself). */
int pcl_cloud_create(const float* xyz_n3_dev, const float* rgb_n3_dev, int64_t n, double q, int order,
                     void* stream, pcl_cloud** out);
int64_t pcl_cloud_size(const pcl_cloud* c);
/* lo_hi[6] = {x_min, y_min, z_min, x_max, y_max, z_max} (host); the first call blocks on the creation stream */
int pcl_cloud_bounds(const pcl_cloud* c, float* lo_hi);
void pcl_cloud_destroy(pcl_cloud* c);

/* ---- equirectangular panorama -------------------------------------------------------------- */
/* img_hw3_dev: (H,W,3) float32 in [0,1]; must stay valid until the work enqueued on `stream` has run (a stream-ordered
 * free is fine).  Formats other than F32 wait once on `stream` for a one-word answer (is the image exactly uint8/255
 * data?), which selects the texel format; the table builds themselves are asynchronous. */
int pcl_image_create(const float* img_hw3_dev, int h, int w, int format, void* stream, pcl_image** out);
int pcl_image_format(const pcl_image* im);
void pcl_image_destroy(pcl_image* im);

/* ---- forward-only scoring of P poses ------------------------------------------------------- */
/* loss_p_dev[p] = Σ m·e / Σ m (NaN when no point survives the zero mask); count_p_dev (nullable) = Σ m */
int pcl_score(const pcl_cloud* c, const pcl_image* im, const float* poses_p6_dev, int64_t p,
              float* loss_p_dev, float* count_p_dev, void* stream);

/* The same table for a structured start grid: every translation i < T against every rotation j < R, output index
 * i*R + j — the double loop of trim_input_loss (utils.py:484-499) over generate_trans_points x generate_rot_points.
 * Rotations of the list that differ by an in-plane turn about the camera z axis (all yaw-only lists; the Euler
 * lattice's 24 rotations = 6 groups of 4) share the rigid transform, elevation and one azimuth atan2 per point.
 * trans_t3_dev: (T,3), rot_r3_dev: (R,3) (yaw, pitch, roll); loss/count: (T*R). */
int pcl_score_grid(const pcl_cloud* c, const pcl_image* im, const float* trans_t3_dev, int64_t t,
                   const float* rot_r3_dev, int r, float* loss_tr_dev, float* count_tr_dev, void* stream);

/* indices of the k smallest losses, ascending, ties -> lower index, NaN last */
int pcl_topk(const float* loss_p_dev, int64_t p, int k, int64_t* idx_k_dev, void* stream);

/* ---- colour-histogram re-rank of K scored candidates (trim_input_hist_secondary, utils.py:510-588) ------- */
/* Renders the cloud from each pose (painter's order of make_pano, utils.py:134-205) and intersects 8x8x8 colour
 * histograms of the middle num_split_h-2 row blocks with the query's.  img_hw3_dev: the (H,W,3) float32 query
 * panorama.  hist_intersect_k_dev[k] = the reference's `hist_intersect` (larger is better). */
int pcl_hist_rerank(const pcl_cloud* c, const float* img_hw3_dev, int h, int w, const float* poses_k6_dev, int k,
                    int num_split_h, int num_split_w, float* hist_intersect_k_dev, void* stream);
/* The same in two stages, for sharding the K candidates over ranks.  Stage 1 is independent per candidate:
 * rows_k_dev[k][2*nblk] (nblk = (num_split_h-2)*num_split_w; per compared block the histogram intersection, then the
 * number of lit rendered pixels) and ngt_dev[nblk] (lit query pixels per block, identical on every rank).  Stage 2 replays
 * the reference's loop over ALL candidates in order (its table persists from one candidate to the next, utils.py:547-579),
 * so it runs on the rows gathered from all ranks.  num_split_h < 3 leaves no compared block: every score is 0 (reference). */
int pcl_hist_rerank_blocks(const pcl_cloud* c, const float* img_hw3_dev, int h, int w, const float* poses_k6_dev, int k,
                           int num_split_h, int num_split_w, float* rows_k_dev, float* ngt_dev, void* stream);
int pcl_hist_rerank_finish(const float* rows_k_dev, const float* ngt_dev, int k, int num_split_h, int num_split_w,
                           float* hist_intersect_k_dev, void* stream);

/* ---- colour matching of the panorama to the cloud (color_match, color_utils.py:146-234; localize.py:402-404) ------ */
/* Both inputs must be uint8/255 data.  pcl_color_stats fills stats_dev (PCL_COLOR_STATS_BYTES, caller-allocated device
 * memory): double whist[3][256] (histogram weighted by row_weight_h_dev[row] = sin(row/H*pi), color_utils.py:216-217, of the lit pixels by TRUNCATED level),
 * uint32 level_cnt[3][256], uint32 value_cnt[3][256] (lit pixels by exact value k), uint64 cloud_cnt[3][256] (points by
 * exact value k), int32 flags[2] (flags[0] != 0: an input was not exactly k/255).  The <= 256-entry cumulative sums and
 * the reference's interpolation are done by the caller (piccolo_b200/color_utils.py); pcl_color_apply rewrites every lit
 * pixel through the resulting look-up table lut[c][k]. */
#define PCL_COLOR_STATS_BYTES (768 * 24 + 16)
int pcl_color_stats(const float* img_hw3_dev, int h, int w, const float* row_weight_h_dev, const float* rgb_n3_dev, int64_t n,
                    void* stats_dev, void* stream);
int pcl_color_apply(const float* img_hw3_dev, int h, int w, const float* lut_3x256_dev, float* out_hw3_dev, void* stream);

/* ---- joint luma equalisation of panorama and cloud (color_mod, color_utils.py:7-65; localize.py:173-179, :405-410) -- */
/* hist_2xbins_dev: uint64 [2][num_bins] luma-bin counts of the lit pixels and of the points (8-bit YCrCb, cv2's fixed-point
 * conversion restated in integers).  The caller forms the cumulative distribution (num_bins floats,
 * piccolo_b200/color_utils.py); pcl_color_mod_apply replaces the luma of every lit pixel and every point by cdf[bin] and
 * converts back, with the reference's uint8 truncations. */
int pcl_color_mod_stats(const float* img_hw3_dev, int h, int w, const float* rgb_n3_dev, int64_t n, int num_bins,
                        unsigned long long* hist_2xbins_dev, void* stream);
int pcl_color_mod_apply(const float* img_hw3_dev, int h, int w, const float* rgb_n3_dev, int64_t n, int num_bins,
                        const float* cdf_bins_dev, float* out_img_hw3_dev, float* out_rgb_n3_dev, void* stream);

/* ---- loss + analytic 6-DoF gradient of B poses (autograd.Function backend) ------------------- */
/* grad_b6_dev[b] = d loss_b / d (tx,ty,tz,yaw,pitch,roll) */
int pcl_loss_fwd_bwd(const pcl_cloud* c, const pcl_image* im, const float* poses_b6_dev, int b,
                     float* loss_b_dev, float* count_b_dev, float* grad_b6_dev, void* stream);

/* ---- fused refinement: one launch per iteration for the whole candidate batch ---------------- */
/* batch_semantics = 0: `omniloc` (forward at the clamped parameter, omniloc.py:44-58)
 * batch_semantics = 1: `omniloc_batch` (forward at the pre-clamp copy, omniloc.py:260-269) */
int pcl_refine_create(int b, double lr, double factor, int patience, int batch_semantics, pcl_refine** out);
/* load B start poses and reset Adam / plateau state */
int pcl_refine_reset(pcl_refine* r, const float* poses_b6_dev, void* stream);
/* run num_iter iterations: each is ONE kernel = loss + backward + reduction + Adam + plateau + clamp; batches of <= 16
 * candidates run ALL iterations in one cooperative launch (cudaLaunchCooperativeKernel; pose blocks alternate so that
 * the per-iteration grid barrier of one block is hidden behind the other block's work), with per-iteration launches of
 * the same arithmetic (bit-identical trajectories) where the device refuses it (option PERSIST=0 forces them) */
int pcl_refine_run(pcl_refine* r, const pcl_cloud* c, const pcl_image* im, int num_iter, void* stream);
/* The same run with the POINTS of the cloud sharded over the ranks of `comm` (the B < #GPUs case: 6 candidates
 * cannot fill 8 GPUs by candidate sharding): every rank holds the same cloud, image and refiner state and evaluates
 * points [N*rank/nranks, N*(rank+1)/nranks); the per-CTA partial sums (14 scalars per candidate) travel as peer stores
 * over NVLink inside the persistent kernel, every rank reduces all records in the same order and steps the optimiser
 * itself, so the states stay bit-identical on all ranks with no host round-trip and no collective library call.
 * All ranks must call with the same B, num_iter and cloud size.  B <= 16. */
int pcl_refine_run_sharded(pcl_refine* r, const pcl_cloud* c, const pcl_image* im, int num_iter, pcl_comm* comm, void* stream);
/* pose_b6_dev: what the reference returns (clamped parameter, or the un-clamped copy under batch
 * semantics); param_b6_dev (nullable): Adam's clamped parameter; loss_b_dev: loss of the LAST forward;
 * lr_b_dev (nullable, double): current learning rates */
int pcl_refine_read(const pcl_refine* r, float* pose_b6_dev, float* param_b6_dev, float* loss_b_dev,
                    double* lr_b_dev, void* stream);
/* Diagnostics (option RF_DEBUG=1 before the run): counters of the last persistent run, 4 per CTA (out_host[cta*4 + k]):
 * compute CTAs (warp 0) {cycles inside phases, cycles waiting for the next poses, phases whose poses were prefetched, 0}; the LAST row is the service CTA {cycles waiting for records, cycles reducing and stepping}.
 * Returns the number of rows recorded (0: nothing recorded) or a negative pcl_status; blocks on `stream`. */
int pcl_refine_debug_stats(const pcl_refine* r, unsigned long long* out_host, int max_ctas, void* stream);
/* Same option: the wall-clock stamp (ns) of the end of every iteration of the last persistent run; returns the count. */
int pcl_refine_debug_timeline(const pcl_refine* r, unsigned long long* out_ns_host, int max_iters, void* stream);
void pcl_refine_destroy(pcl_refine* r);

/* ---- peer-memory communicator of the ranks of one box (one process per GPU) ---------------------------------- */
/* The reference is single-device (localize.py:124); SURVEY 8b/8e ask for the sharding glue in the ABI so that a caller
 * without torch.distributed can shard.  Every rank creates a window (cudaMalloc, zero-filled; window_bytes 0 = 16 MB),
 * the 64-byte CUDA IPC handles are exchanged through any host channel, pcl_comm_connect maps the peers' windows, and
 * from then on kernels exchange data by stores and system-scope atomics over NVLink.  All ranks must issue the same
 * sequence of collective calls (barrier, allgather, sharded refinement). */
#define PCL_COMM_HANDLE_BYTES 64
int pcl_comm_create(int rank, int nranks, size_t window_bytes, pcl_comm** out);
int pcl_comm_handle(const pcl_comm* c, void* handle64);
int pcl_comm_connect(pcl_comm* c, const void* handles_nranks_x_64);
/* alternative: peer windows mapped by the caller (e.g. torch symmetric memory), zero-filled, >= window_bytes each */
int pcl_comm_connect_ptrs(pcl_comm* c, void* const* windows);
int pcl_comm_rank(const pcl_comm* c);
int pcl_comm_size(const pcl_comm* c);
/* stream-ordered barrier: work enqueued after it on `stream` starts when every rank's earlier work has finished */
int pcl_comm_barrier(pcl_comm* c, void* stream);
/* dst_dev[k*n + i] = src_dev[i] of rank k, n <= 16384 floats per rank — the all-gather of per-pose losses after sharded
 * scoring and of the (loss, pose) rows before the arg-min */
int pcl_comm_allgather_f32(pcl_comm* c, const float* src_dev, int n, float* dst_dev, void* stream);
void pcl_comm_destroy(pcl_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* PICCOLO_B200_H */
