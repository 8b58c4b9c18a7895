/*
 * dsrg_b200.h -- C ABI of the B200-native DSRG pixel-labelling hot path.
 *
 * Plain C, no torch / C++ types.  All `*_dev` pointers are CUDA device pointers on the
 * engine's device (e.g. torch.Tensor.data_ptr()); all `*_host` pointers are host memory
 * (pinned memory from dsrg_host_alloc() makes the copies asynchronous).  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Every function returning int
 * returns 0 on success and a negative DSRG_E_* code on failure; dsrg_last_error() then holds a
 * human-readable message (thread-local).  Nothing in this library falls back to a CPU path:
 * without a usable sm_100 device every call fails with DSRG_E_CUDA.
 *
 * Each entry point names the reference interface (speedinghzl/DSRG, path:line) it replaces.
 * Layout names: NHWC = [B][H][W][M] (pixel-major, what krahenbuhl2013.CRF and DenseCRFWrapper
 * use), NCHW = [B][M][H][W] (Caffe blobs, what pylayers.py passes around).
 */
#ifndef DSRG_B200_H
#define DSRG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSRG_OK 0
#define DSRG_E_INVALID (-1)   /* bad argument (shape, layout, NULL pointer, batch > max_batch) */
#define DSRG_E_CUDA (-2)      /* CUDA runtime error, or no sm_100 device                        */
#define DSRG_E_KEYRANGE (-3)  /* lattice coordinates exceed the packed-key range */
#define DSRG_E_STATE (-4)     /* call order violated (e.g. inference before add_pairwise)       */
#define DSRG_E_NOMEM (-5)

#define DSRG_LAYOUT_NHWC 0
#define DSRG_LAYOUT_NCHW 1

#define DSRG_MAX_LABELS 32 /* labels per pixel of the fused, register-tiled kernels (the 21-class hot path) */
/* CRF (batch, layer and DenseCRF-object entry points) and SRG accept up to this many labels: above DSRG_MAX_LABELS
 * the mean field runs on a generic label-chunked path (csrc/meanfield_wide.cu) -- the reference's COCO tool uses
 * DenseCRF(W, H, 81), training/tools/test-coco.py.  The Softmax / ConstrainLoss layers and the predict_mask
 * post-processing are built for at most DSRG_MAX_LABELS. */
#define DSRG_MAX_LABELS_WIDE 255

int dsrg_version(void);               /* 10000*major + 100*minor + patch */
const char *dsrg_last_error(void);    /* message of the last failing call on this thread */
int dsrg_device_count(void);          /* number of visible CUDA devices (0 if none / no driver) */
/* The device an engine is created on when `device` is -1, and the one the drop-ins (pylayers, krahenbuhl2013,
 * the DenseCRF objects) use: DSRG_B200_DEVICE if set, else the calling thread's current CUDA device -- what
 * caffe.set_device(N) (training/tools/train.py:77-79) or torch.cuda.set_device selected.  -1 if there is none.
 * Every entry point restores the caller's current device before it returns. */
int dsrg_current_device(void);

/* Pinned host memory so the *_host entry points overlap their copies (cudaHostAlloc). */
void *dsrg_host_alloc(size_t bytes);
void dsrg_host_free(void *p);
/* Page-lock memory the caller already owns (a Caffe blob's CPU buffer in CPU mode, a numpy array) so that the
 * *_host entry points copy from / to it asynchronously at full PCIe rate (cudaHostRegister / Unregister). */
int dsrg_host_register(void *p, size_t bytes);
int dsrg_host_unregister(void *p);

/* ------------------------------------------------------------------------------------------
 * Pairwise parameters of krahenbuhl2013.CRF -- CRF/krahenbuhl2013/CRF.py:31-32 passes
 * (w1=10, 80/s, 80/s, 13, 13, 13, w2=3, 3/s, 3/s); argument order and meaning are those of
 * DenseCRFWrapper::add_pairwise_energy, CRF/src/densecrf_wrapper.cpp:18-30
 * (Gaussian/spatial kernel w2 is applied BEFORE the bilateral kernel w1).
 * ------------------------------------------------------------------------------------------ */
typedef struct dsrg_crf_params {
    float w1;            /* bilateral Potts weight                     */
    float theta_alpha_x; /* bilateral spatial sigma (x)                */
    float theta_alpha_y; /* bilateral spatial sigma (y)                */
    float theta_beta_r;  /* bilateral colour sigmas, channel 0 / 1 / 2 */
    float theta_beta_g;
    float theta_beta_b;
    float w2;            /* spatial Potts weight                       */
    float theta_gamma_x; /* spatial sigma (x)                          */
    float theta_gamma_y; /* spatial sigma (y)                          */
    int n_iters;         /* mean-field iterations (CRF.py:4 maxiter=10) */
} dsrg_crf_params;

/* Fill `p` exactly as CRF.py:31-32 does for a given scale_factor / color_factor / maxiter. */
void dsrg_crf_params_default(dsrg_crf_params *p, float scale_factor, float color_factor, int maxiter);

/* ------------------------------------------------------------------------------------------
 * Engine: owns every device buffer the batched kernels need for up to max_batch images of
 * H x W pixels and M labels on `device` (-1: dsrg_current_device()).  No allocation happens on the hot calls.
 * ------------------------------------------------------------------------------------------ */
typedef struct dsrg_engine dsrg_engine;

dsrg_engine *dsrg_engine_create(int device, int max_batch, int H, int W, int M);
void dsrg_engine_destroy(dsrg_engine *e);
size_t dsrg_engine_device_bytes(const dsrg_engine *e); /* bytes of HBM held by the engine */
/* The H x W given to dsrg_engine_create is a capacity: any H' <= H, W' <= W can be selected afterwards
 * without reallocating (the evaluation tools run one image at a time, each of its own size --
 * training/tools/test-ms.py:86-87).  Waits for queued work; the cached spatial lattice is rebuilt. */
int dsrg_engine_set_size(dsrg_engine *e, int H, int W);
int dsrg_engine_get_size(const dsrg_engine *e, int *H, int *W, int *H_capacity, int *W_capacity);
/* The *_host full-pass entry points pipeline the batch in chunks (default B/8, 3B/8, B/2 images) through
 * H2D | kernels | D2H streams; `images` > 0 caps the chunk size, 0 restores the default. */
int dsrg_engine_set_host_chunk(dsrg_engine *e, int images);
/* Experimental: run the mean-field loop as `lanes` (1 or 2, default 1) half-batches on separate
 * streams so that the DRAM-latency-bound blur passes of one half overlap the shared-memory-bound tile
 * kernel of the other (measured on B200: +1 %, so it is off by default). */
int dsrg_engine_set_lanes(dsrg_engine *e, int lanes);
/* Device passes (*_dev entry points, and through them the chunks of the *_host ones) are captured into CUDA
 * graphs the second time a pass is issued with the same arguments and replayed afterwards -- one launch instead
 * of 40-130 dependent ones.  Needs a real stream (not the legacy default stream); enable = 0 turns it off and
 * drops the cached graphs (also: environment DSRG_B200_GRAPHS=0).  graph_replays counts the passes replayed. */
int dsrg_engine_set_graphs(dsrg_engine *e, int enable);
long long dsrg_engine_graph_replays(const dsrg_engine *e);
/* Kernel launches issued by this engine since the last call (bench.py's gpu_launches); the kernels inside a
 * replayed graph are counted. */
long long dsrg_engine_take_launch_count(dsrg_engine *e);
/* Diagnostic: how many tiles of the last mean-field pass took the hybrid path (the shared-memory list of the tile's
 * most-touched bilateral vertices plus direct global slices / reductions for the rest -- textured images; see
 * csrc/tiles.cu).  Synchronises the device; < 0 on error.  No reference analogue (permutohedral.cpp:545-584 visits
 * every (pixel, vertex) incidence the same way). */
long long dsrg_engine_hybrid_tiles(dsrg_engine *e);

/*
 * Batched dense-CRF mean-field inference; replaces the per-image loop
 *   for i in range(N): result[i] = CRF(im[i], unary[i], scale_factor)
 * of pylayers/pylayers/pylayers.py:81-82 / :325-326, i.e. B x { CRF.py:4-37 ->
 * wrapper.pyx:23-60 -> densecrf_wrapper.cpp:5-50 -> DenseCRF::inference, densecrf.cpp:115-131 }.
 *   unary : B x (H,W,M) values as passed to CRF(image, unary): energy = -unary (CRF.py:28)
 *   image : B x (H,W,3) uint8 (CRF.py:32 casts to ubyte)
 *   out   : B x marginals Q, float32
 * Results are within 1e-4 (max abs) of the reference; see DESIGN.md for why not bit-exact.
 */
int dsrg_crf_batch_dev(dsrg_engine *e, int B, const float *unary_dev, int unary_layout,
                       const uint8_t *image_dev, const dsrg_crf_params *params, float *out_dev,
                       int out_layout, void *stream);
int dsrg_crf_batch_host(dsrg_engine *e, int B, const float *unary_host, int unary_layout,
                        const uint8_t *image_host, const dsrg_crf_params *params, float *out_host,
                        int out_layout);
/* arg-max labelling after inference: DenseCRFWrapper::map, densecrf_wrapper.cpp:39-43 */
int dsrg_crf_map_batch_dev(dsrg_engine *e, int B, const float *unary_dev, int unary_layout,
                           const uint8_t *image_dev, const dsrg_crf_params *params,
                           int32_t *labels_out_dev, void *stream);

/*
 * Batched seeded region growing; replaces
 *   self.pool.map(generate_seed_step, items)     pylayers/pylayers/pylayers.py:341-342
 * = B x generate_seed_step (pylayers.py:237-275) incl. CC_labeling_8.CC_lab
 * (pylayers/pylayers/CC_labeling_8.py:103-282).
 *   labels : [B][M]       image-level tags, class c present iff labels[c] == 1
 *   probs  : [B][M][H][W] float32 (compared in float64 like the reference's data)
 *   cues   : [B][M][H][W] float32 0/1 seeds
 *   th1,th2: background / foreground thresholds as float64 (the reference compares Python
 *            floats 0.99 / 0.85 against float64 data with strict >)
 *   renorm : 1 = first apply the post-CRF clamp(1e-4) + float64 renormalisation of
 *            pylayers.py:328-330 to `probs` (the DSRGLayer path); 0 = use probs as given
 *   seeds_out     : [B][M][H][W] float32 0/1 (old seeds are kept, pylayers.py:275)
 *   label_map_out : optional [B][H][W] int32 (0 = none, c+1 = class c), may be NULL
 * Bit-exact against the reference for identical inputs.
 */
int dsrg_srg_batch_dev(dsrg_engine *e, int B, const float *labels_dev, const float *probs_dev,
                       const float *cues_dev, double th1, double th2, int renorm,
                       float *seeds_out_dev, int32_t *label_map_out_dev, void *stream);
int dsrg_srg_batch_host(dsrg_engine *e, int B, const float *labels_host, const float *probs_host,
                        const float *cues_host, double th1, double th2, int renorm,
                        float *seeds_out_host, int32_t *label_map_out_host);

/*
 * The whole DSRGLayer.forward body (pylayers.py:297-304, :333-344): refinement
 * (pylayers.py:310-331: in-place clamp of probs at 1e-4, CRF with unary = probs, clamp +
 * float64 renormalise) followed by SRG.  `image` is the already zoomed / mean-added / rounded
 * uint8 image at the probs resolution (the host shim does pylayers.py:315-319).
 *   probs_dev     : [B][M][H][W], CLAMPED IN PLACE like the reference's bottom blob (:312)
 *   crf_out_dev   : optional [B][M][H][W] float32 RAW CRF marginals of this very call, i.e. before
 *                   the clamp + float64 renormalisation of :328-330 (NULL to skip)
 */
int dsrg_dsrg_forward_dev(dsrg_engine *e, int B, const float *labels_dev, float *probs_dev,
                          const float *cues_dev, const uint8_t *image_dev,
                          const dsrg_crf_params *params, double th1, double th2,
                          float *seeds_out_dev, float *crf_out_dev, void *stream);
int dsrg_dsrg_forward_host(dsrg_engine *e, int B, const float *labels_host, float *probs_host,
                           const float *cues_host, const uint8_t *image_host,
                           const dsrg_crf_params *params, double th1, double th2,
                           float *seeds_out_host, float *crf_out_host);

/*
 * SURVEY 8f rank 1 -- the producer of `probs` and the other consumer of the CRF result:
 * SoftmaxLayer (pylayers.py:23-51): probs = (softmax(preds) + 1e-4) / sum(...); backward = d sum(probs*top_diff)/d preds.
 * ConstrainLossLayer (pylayers.py:154-180): loss = mean_{n,h,w} sum_c ps log(clip(ps/probs, 0.05, 20)), ps = exp(log_smooth);
 * backward writes both gradients (pylayers.py:176-180).  All arrays [B][M][H][W] float32.
 */
int dsrg_softmax_forward_dev(dsrg_engine *e, int B, const float *preds_dev, float *probs_out_dev, void *stream);
int dsrg_softmax_backward_dev(dsrg_engine *e, int B, const float *preds_dev, const float *top_diff_dev,
                              float *grad_out_dev, void *stream);
int dsrg_constrainloss_forward_dev(dsrg_engine *e, int B, const float *probs_dev, const float *log_smooth_dev,
                                   float *loss_out_dev /* 1 float */, void *stream);
int dsrg_constrainloss_backward_dev(dsrg_engine *e, int B, const float *probs_dev, const float *log_smooth_dev,
                                    float *grad_probs_dev, float *grad_log_dev, void *stream);
int dsrg_softmax_forward_host(dsrg_engine *e, int B, const float *preds_host, float *probs_out_host);
int dsrg_softmax_backward_host(dsrg_engine *e, int B, const float *preds_host, const float *top_diff_host,
                               float *grad_out_host);
int dsrg_constrainloss_forward_host(dsrg_engine *e, int B, const float *probs_host, const float *log_smooth_host,
                                    float *loss_out_host /* 1 float */);
int dsrg_constrainloss_backward_host(dsrg_engine *e, int B, const float *probs_host, const float *log_smooth_host,
                                     float *grad_probs_host, float *grad_log_host);

/*
 * Image preprocessing of CRFLayer / DSRGLayer (pylayers.py:70-75, :315-319) + the ubyte cast of
 * CRF.py:32: bilinear zoom (scipy.ndimage.zoom order=1 semantics, float64 arithmetic, bit-exact) of the
 * [B][3][Hi][Wi] float32 network input to the engine's H x W, + mean_pixel[3], round half to even, ->
 * [B][H][W][3] uint8, the `image` argument of the entry points above.
 */
int dsrg_prepare_image_dev(dsrg_engine *e, int B, int Hi, int Wi, const float *images_dev,
                           const double *mean_pixel /* host, 3 */, uint8_t *image_out_dev, void *stream);
int dsrg_prepare_image_host(dsrg_engine *e, int B, int Hi, int Wi, const float *images_host,
                            const double *mean_pixel /* host, 3 */, uint8_t *image_out_host);

/*
 * Full-resolution inference post-processing: what predict_mask() of the evaluation / ground-truth tools
 * does between the network's score blob and the label map (engine batch 1, size = the image's):
 *   DSRG_POST_SUM_SCORES  training/tools/test-ms.py:84-111: scores_all = sum_k zoom(scores_k, order=1);
 *                         softmax over labels; clamp at eps; CRF(im, log(probs)); argmax
 *   DSRG_POST_ZOOM_PROBS  training/tools/generate_train_gt.py:76-104: softmax at network resolution;
 *                         zoom(probs, order=1); clamp at eps; CRF(im, log(probs)); argmax over `labels_sel`
 *   scores     : n_scales pointers to [M][h_k][w_k] float32 (net.blobs['fc8-SEC'].data[0]); the array of
 *                pointers and hs / ws live on the host in both variants
 *   image      : [H][W][3] uint8 as handed to krahenbuhl2013.CRF (may be NULL when smooth == 0)
 *   smooth     : 0 skips the CRF (the tools' `smooth=False`)
 *   labels_sel : n_sel label ids ([0] + image tags in generate_train_gt.py:96-97); n_sel == 0 = all labels
 *   result_out : [H][W] int32 label map;  probs_out : optional [H][W][M] float32 (CRF marginals, or the
 *                clamped probabilities when smooth == 0)
 * dsrg_zoom_scores_* is the zoom step alone ([M][h][w] -> [H][W][M], scipy.ndimage.zoom order=1 semantics,
 * bit-exact; accumulate != 0 adds to `out` in float32 like `scores_all += scores`).
 */
#define DSRG_POST_SUM_SCORES 0
#define DSRG_POST_ZOOM_PROBS 1
int dsrg_zoom_scores_dev(dsrg_engine *e, const float *scores_dev, int h, int w, float *out_dev, int accumulate,
                         void *stream);
int dsrg_zoom_scores_host(dsrg_engine *e, const float *scores_host, int h, int w, float *out_host,
                          int accumulate);
int dsrg_predict_mask_dev(dsrg_engine *e, int mode, int n_scales, const float *const *scores_dev, const int *hs,
                          const int *ws, const uint8_t *image_dev, float eps, int smooth,
                          const dsrg_crf_params *params, const int32_t *labels_sel, int n_sel,
                          int32_t *result_out_dev, float *probs_out_dev, void *stream);
int dsrg_predict_mask_host(dsrg_engine *e, int mode, int n_scales, const float *const *scores_host,
                           const int *hs, const int *ws, const uint8_t *image_host, float eps, int smooth,
                           const dsrg_crf_params *params, const int32_t *labels_sel, int n_sel,
                           int32_t *result_out_host, float *probs_out_host);

/*
 * AnnotationLayer.forward (pylayers/pylayers/pylayers.py:369-387): image tags + sparse localisation cues
 * -> the dense `labels` [B][1][1][M] and `cues` [B][M][H][W] blobs (H x W = the engine's size), and the
 * optionally mirrored copy of the images.  Reading the pickle and drawing `flip` stay with the caller:
 *   tag_offsets [B+1], tags       : CSR of data_file['%i_labels'] per image (class ids; label 0 is always set)
 *   cue_offsets [B+1], cue_idx    : CSR of data_file['%i_cues']; cue_idx is [3][K] (class, row, column rows,
 *                                   K = cue_offsets[B]); negative indices wrap like numpy's, others are errors
 *   flip [B] or NULL              : 1 = mirror this image's cues and pixels along the width (flip == -1, :385)
 *   images_in / images_out        : [B][3][Hi][Wi] float32, distinct buffers; both NULL = skip the copy
 * All index arrays live on the host in both variants.
 */
int dsrg_annotation_forward_dev(dsrg_engine *e, int B, const int32_t *tag_offsets, const int32_t *tags,
                                const int32_t *cue_offsets, const int32_t *cue_idx, const int32_t *flip,
                                const float *images_in_dev, int Hi, int Wi, float *labels_out_dev,
                                float *cues_out_dev, float *images_out_dev, void *stream);
int dsrg_annotation_forward_host(dsrg_engine *e, int B, const int32_t *tag_offsets, const int32_t *tags,
                                 const int32_t *cue_offsets, const int32_t *cue_idx, const int32_t *flip,
                                 const float *images_in_host, int Hi, int Wi, float *labels_out_host,
                                 float *cues_out_host, float *images_out_host);

/* Host-only helpers of the *_host wire format (0/1 planes cross PCIe as 1 bit per value, see
 * csrc/wire.cu); exported for unit tests.  pack returns 1 if every value was exactly 0 or 1. */
int dsrg_wire_pack_mask(const float *src_host, uint32_t *dst_bits, size_t n);
void dsrg_wire_unpack_mask(const uint32_t *src_bits, float *dst_host, size_t n);
void dsrg_wire_apply_clamp_mask(const uint32_t *src_bits, float *probs_host, size_t n);

/*
 * CRFLayer.forward body (pylayers.py:63-88): same refinement, output log(result).
 *   log_out_dev : [B][M][H][W] float32 = log(renormalised marginals)
 *   result_dev  : optional [B][M][H][W] float32 copy of the marginals kept for backward (:90-92)
 */
int dsrg_crflayer_forward_dev(dsrg_engine *e, int B, float *probs_dev, const uint8_t *image_dev,
                              const dsrg_crf_params *params, float *log_out_dev,
                              float *result_dev, void *stream);

int dsrg_crflayer_forward_host(dsrg_engine *e, int B, float *probs_host, const uint8_t *image_host,
                               const dsrg_crf_params *params, float *log_out_host,
                               float *result_host);

/*
 * One refinement, two consumers.  In the reference's net CRFLayer and DSRGLayer are fed the same two blobs
 * (train-s.prototxt:758-786) and each runs the whole dense CRF on them (pylayers.py:82 and :326).  The raw
 * marginals of an engine's last mean-field pass stay on the device; these entry points let a second consumer
 * use them instead of repeating the pass:
 *   dsrg_srg_last_crf_host       : generate_seed_step over the batch on those marginals (float64 clamp +
 *                                  renormalisation fused in, exactly what dsrg_dsrg_forward_* does after its CRF)
 *   dsrg_crf_last_marginals_host : the raw float32 marginals themselves
 * Both return DSRG_E_STATE unless the engine's last pass was a CRF over exactly B images.
 */
int dsrg_srg_last_crf_host(dsrg_engine *e, int B, const float *labels_host, const float *cues_host, double th1,
                           double th2, float *seeds_out_host);
int dsrg_crf_last_marginals_host(dsrg_engine *e, int B, float *out_host, int out_layout);

/*
 * BalancedSeedLossLayer (pylayers.py:120-152).  Forward writes the LOCAL sums
 *   terms_out[0] = sum_n S_bg(n) / max(cnt_bg(n), 1e-4),  terms_out[1] = same for fg
 * so that loss = -(terms[0] + terms[1]) / N_global; with several GPUs the two floats are
 * all-reduced (SUM) first -- the only collective on the path.  Backward writes
 *   grad = top_diff * d loss / d probs = -top_diff * lab / (p * max(cnt,1e-4) * N_global).
 */
int dsrg_seedloss_forward_dev(dsrg_engine *e, int B, const float *probs_dev,
                              const float *seeds_dev, float *terms_out_dev, void *stream);
int dsrg_seedloss_backward_dev(dsrg_engine *e, int B, int n_global, const float *probs_dev,
                               const float *seeds_dev, float top_diff, float *grad_out_dev,
                               void *stream);

int dsrg_seedloss_forward_host(dsrg_engine *e, int B, const float *probs_host,
                               const float *seeds_host, float *terms_out_host /* 2 floats */);
int dsrg_seedloss_backward_host(dsrg_engine *e, int B, int n_global, const float *probs_host,
                                const float *seeds_host, float top_diff, float *grad_out_host);

/* Optional per-kernel timing for the roofline report: while enabled every kernel launch of the
 * engine is bracketed by CUDA events on its launching stream.  dsrg_engine_profile_read()
 * synchronises the device and returns, per kernel class (dsrg_profile_tag_count() of them, named by
 * dsrg_profile_tag_name()), the summed milliseconds and the number of launches since the last read. */
int dsrg_profile_tag_count(void);
const char *dsrg_profile_tag_name(int tag);
int dsrg_engine_profile(dsrg_engine *e, int enable);
int dsrg_engine_profile_read(dsrg_engine *e, float *ms_out, long long *count_out);

/* Introspection used by the lattice-level parity tests: vertex counts of the lattices built by
 * the last CRF call (spatial, then bilateral per image). */
int dsrg_engine_lattice_sizes(dsrg_engine *e, int B, int *v_spatial_out, int *v_bilateral_out);
/* Per-pixel symmetric normalisation vectors (DenseKernel::norm_, pairwise.cpp:54-57) of the
 * last CRF call: which = 0 spatial [N] (shared by the batch), 1 bilateral [B][N]. */
int dsrg_engine_copy_norm(dsrg_engine *e, int which, int B, float *norm_out_host);

/* ------------------------------------------------------------------------------------------
 * Per-object API with the exact surface of the reference's C++ class DenseCRFWrapper
 * (CRF/include/densecrf_wrapper.h:3-28), which krahenbuhl2013/wrapper.pyx:20-60 binds.
 * Host pointers, borrowed for the duration of the call, like the original.  Creating an object is cheap
 * (the reference makes one per image, CRF.py:21): it allocates nothing on the device.
 * ------------------------------------------------------------------------------------------ */
typedef struct dsrg_densecrf dsrg_densecrf;

dsrg_densecrf *dsrg_densecrf_create(int W, int H, int nlabels);            /* densecrf_wrapper.cpp:5-8   */
void dsrg_densecrf_destroy(dsrg_densecrf *c);                              /* :10-12                     */
int dsrg_densecrf_npixels(const dsrg_densecrf *c);                         /* :14                        */
int dsrg_densecrf_nlabels(const dsrg_densecrf *c);                         /* :15                        */
int dsrg_densecrf_set_unary_energy(dsrg_densecrf *c, const float *unary_costs); /* :32-37, [N][M] energies */
int dsrg_densecrf_add_pairwise_energy(dsrg_densecrf *c, float w1, float theta_alpha_1,
                                      float theta_alpha_2, float theta_betta_1, float theta_betta_2,
                                      float theta_betta_3, float w2, float theta_gamma_1,
                                      float theta_gamma_2, const unsigned char *im);  /* :18-30 */
int dsrg_densecrf_map(dsrg_densecrf *c, int n_iters, int *labels);         /* :39-43 */
int dsrg_densecrf_inference(dsrg_densecrf *c, int n_iters, float *probs_out); /* :45-50 */
/* The objects borrow a process-wide engine per label count (created on first use, grown to the largest image
 * seen, serialised by a mutex); this frees those engines' device memory. */
void dsrg_densecrf_release_engines(void);

#ifdef __cplusplus
}
#endif
#endif /* DSRG_B200_H */
