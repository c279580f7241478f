/*
 * jolineedle_b200 -- C ABI of the B200-native gaze-environment hot path.
 *
 * The reference (jolibrain/jolineedle) is pure Python: it has no plugin / operator / FFI
 * boundary of its own.  Each entry point below therefore cites the reference *Python*
 * function whose work it replaces (paths relative to the reference root).  All pointers are
 * plain device pointers unless marked HOST; the library borrows them for the duration of the
 * call and owns nothing but the opaque handles it hands out.  Every launch goes to the
 * `stream` argument (a cudaStream_t passed as void*; NULL = legacy default stream).
 *
 * Return value: 0 (JN_OK) on success, a jn_status code otherwise; jn_last_error() gives a
 * thread-local, human-readable description of the last failure.
 *
 * Conventions shared by all entry points
 *   - positions are int64 [n, 2] in (row, col) = (y, x) PATCH coordinates (torch.long, as in
 *     the reference: general_env.py:118-123);
 *   - action codes are the reference's (src/env/common.py:4-15): 0 LEFT, 1 RIGHT, 2 UP,
 *     3 DOWN, 4 LEFT_UP, 5 RIGHT_UP, 6 LEFT_DOWN, 7 RIGHT_DOWN, 8 STOP;
 *   - boxes are int64 [.., 4] in x1, y1, x2, y2 PIXEL coordinates (src/utils.py:95-106);
 *   - patch bitmaps are uint32 words, bit (y * cols + x) & 31 of word (y * cols + x) >> 5,
 *     `jn_bitmap_words(rows, cols)` words per episode;
 *   - bool tensors are one byte per element holding 0 or 1 (torch.bool storage).
 */
#ifndef JOLINEEDLE_B200_H
#define JOLINEEDLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JN_ABI_VERSION 1

typedef enum jn_status {
  JN_OK = 0,
  JN_ERR_INVALID = 1,     /* bad argument (maps to AssertionError / ValueError on the Python side) */
  JN_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed */
  JN_ERR_UNSUPPORTED = 3, /* valid request that this build does not cover */
  JN_ERR_NO_DEVICE = 4    /* no sm_100 device visible */
} jn_status;

typedef enum jn_dtype { JN_U8 = 0, JN_F32 = 1 } jn_dtype;

/* jn_gather flags */
#define JN_GATHER_NORMALIZE 1u /* uint8 source -> float32 output, value / 255 (ToTensor, dataset.py:240) */
#define JN_GATHER_FOCUS 2u     /* output in YOLOX Focus space-to-depth layout [4*C, P/2, P/2] */
#define JN_GATHER_SHIFT_ALIGNED 4u /* caller guarantees every x shift is a multiple of 16 bytes (TMA may serve it) */
#define JN_GATHER_SKIP_NEGATIVE 8u /* items with a negative src_index are left untouched instead of zero-filled */
/* src_index values: >= 0 image; -1 zero-filled tile (left untouched with JN_GATHER_SKIP_NEGATIVE);
 * <= JN_SRC_SKIP tile always left untouched */
#define JN_SRC_SKIP (-2)

/* jn_gather engine selection (JN_ENGINE_AUTO picks the fastest one the shapes allow) */
typedef enum jn_engine {
  JN_ENGINE_AUTO = 0,
  JN_ENGINE_TENSOR = 1, /* TMA tensor-map tiles  (cp.async.bulk.tensor, SASS UTMALDG)  */
  JN_ENGINE_BULK = 2,   /* TMA row copies        (cp.async.bulk,        SASS UBLKCP)   */
  JN_ENGINE_LDG = 3     /* plain global loads: any alignment, any stride (slow, always valid) */
} jn_engine;

/* overlap rules of jn_patch_bitmaps */
typedef enum jn_overlap_rule {
  JN_RULE_ANY_PIXEL = 0, /* NeedleGeneralEnv.convert_bboxes_to_masks, general_env.py:360-379 */
  JN_RULE_AREA5 = 1      /* NeedleSimpleEnv.bbox_positions (>5% of P^2 or centre), simple_env.py:270-321 */
} jn_overlap_rule;

int jn_abi_version(void);
const char* jn_last_error(void);
/* Kernels launched by the library since it was loaded (all entry points, all threads). */
long long jn_launch_count(void);
/* First 16 hex digits of the SHA-256 of the native sources this library was built from (csrc/Makefile);
 * profiles are stamped with it. */
const char* jn_source_hash(void);
/* SM count and compute capability of the current device. */
int jn_device_info(int* sm_count, int* cc_major, int* cc_minor);
static inline int jn_bitmap_words(int rows, int cols) { return (rows * cols + 31) / 32; }
/* Host-only self test of arithmetic shared with the kernels (runs without a GPU):
 * unit_out HOST float[256] <- the uint8 -> value/255 normalisation of every byte value;
 * direction_out HOST int[9] <- action code for gradient signs, index (sign(dy)+1)*3 + sign(dx)+1
 * (the decision table of move_towards, simple_env.py:84-125). */
int jn_selftest_host(float* unit_out, int* direction_out);

/* ------------------------------------------------------------------------------------------
 * Image sets: where glimpses are gathered from.
 *
 * An image set describes `n_slabs` device allocations, slab k holding `counts[k]` images of
 * shape [C, heights[k], widths[k]] back to back (a [B, C, H, W] batch is one slab; a python
 * list of [C, H, W] tensors is B slabs of one image).  Images are numbered in slab order.
 * Creating a set encodes the TMA tensor maps for it.  Replaces the `self.images` /
 * `self.image` members of the reference envs (general_env.py:84-115 with n_glimps_levels=1,
 * simple_env.py:184) -- no pixel is copied.
 * ------------------------------------------------------------------------------------------ */
typedef struct jn_images jn_images;

int jn_images_create(jn_images** out, int n_slabs, const void* const* slab_ptrs /*HOST*/,
                     const int32_t* counts /*HOST*/, const int32_t* heights /*HOST*/,
                     const int32_t* widths /*HOST*/, int channels, int dtype /*jn_dtype*/,
                     int patch_size, void* table_host /*HOST*/, void* table_dev, void* stream);
/* Multi-slab sets keep a per-image table (jn_images_table_bytes(total images) bytes) in device
 * memory.  With `table_host` and `table_dev` both given, the library only WRITES the table into
 * the caller's host scratch `table_host`; the caller copies it to `table_dev` on its stream before
 * the first gather and keeps both alive as long as the set (no CUDA memory call is made, nothing
 * synchronises).  With both NULL the library allocates and uploads the table itself
 * (cudaMalloc / cudaFree synchronise the device). */
static inline int64_t jn_images_table_bytes(int64_t n_images) { return n_images * 32; }
/* Same, for images whose height / width are NOT multiples of patch_size: the set behaves as if every
 * image had been zero-padded at the bottom / right up to the next multiple (complete_to_patch_size and
 * padded_collate_fn, dataset.py:307-347,379-406) -- the patch grid is ceil(H/P) x ceil(W/P) and the pixels
 * of the edge tiles that lie outside the image are zeros -- without materialising the padding (the TMA unit
 * zero-fills out-of-bounds bytes; sets the tensor engine cannot address fall back to plain loads).  With
 * `shifts`, the image is translated inside its own H x W frame first and padded afterwards (the dataset's order:
 * transform, then collate), i.e. tile pixels whose frame coordinate lies in the padding are zeros. */
int jn_images_create_padded(jn_images** out, int n_slabs, const void* const* slab_ptrs /*HOST*/,
                            const int32_t* counts /*HOST*/, const int32_t* heights /*HOST*/,
                            const int32_t* widths /*HOST*/, int channels, int dtype /*jn_dtype*/,
                            int patch_size, void* table_host /*HOST*/, void* table_dev, void* stream);
void jn_images_destroy(jn_images* set);
/* 1 if the TMA engines can serve this set (16-byte aligned bases / rows / patches), else 0. */
int jn_images_tma_ok(const jn_images* set, int engine /*jn_engine*/);

/* ------------------------------------------------------------------------------------------
 * K1  glimpse gather.
 *
 * For item i in [0, n_items): image = src_index ? src_index[i] : i (negative -> the output
 * tile is zero-filled), (y, x) = positions[i]; copies the [C, P, P] tile at patch (y, x) of
 * that image to out + i * out_item_stride_bytes.  Output dtype = source dtype, or float32
 * with JN_GATHER_NORMALIZE; layout [C, P, P], or [4C, P/2, P/2] with JN_GATHER_FOCUS
 * (out[(dy + 2*dx) * C + c][i][j] = tile[c][2i + dy][2j + dx], the YOLOX Focus stem order
 * TL, BL, TR, BR).
 *
 * Replaces: NeedleGeneralEnv.patches (general_env.py:285-306), get_patch (simple_env.py:55-81)
 * plus the per-step copy into the sample (simple_env.py:472), the patch loop of
 * get_detection_batch (general_env.py:531-542) and of init_sample (simple_env.py:417-419).
 *
 * `shifts` (device int32 [n_images, 2] = (ty, tx) per image, or NULL): the tiles are taken from the
 * image translated by (tx, ty) pixels with zero fill -- tile pixel (r, c) of patch (y, x) is image
 * pixel (y*P + r - ty, x*P + c - tx) -- i.e. the integer `translate` augmentation of the reference's
 * dataset (dataset.py:157-226, torchvision F.affine with fill 0) folded into the gather instead of
 * materialising a shifted copy of the image.  One slab: any offset rides the TMA engine (the unit only
 * accepts inner offsets that are 16-byte multiples, so the converting kernel loads the aligned superset of
 * every row and realigns it in shared memory; same-dtype uint8 copies need JN_GATHER_SHIFT_ALIGNED and
 * P <= 256).  Lists of images: plain-load kernel (one warp per tile row).
 *
 * `positions` may be NULL: patch (0, 0) of every item's image (sets of one-patch images, e.g. a crop
 * history registered as an image set).
 *
 * `status` (device int32[1], may be NULL) is OR-ed with 1 if some position was outside the
 * patch grid (that tile is skipped).
 * ------------------------------------------------------------------------------------------ */
int jn_gather(const jn_images* set, const int64_t* positions, const int32_t* src_index,
              const int32_t* shifts, int n_items, void* out, int64_t out_item_stride_bytes,
              uint32_t flags, int engine /*jn_engine*/, int32_t* status, void* stream);

/* Host-only: the ticket schedule the converting gather uses for a launch of `total_chunks` chunks on `grid`
 * CTAs (work is claimed from a global counter in batches of `batch` chunks -- a power of two <= 32 -- shrinking
 * to single chunks at the end of the launch).  Returns the number of segments n (<= 6); segment j hands out tickets [tickets[j], tickets[j+1])
 * as batches of sizes[j] chunks starting at chunks[j].  For tests: the tickets tile [0, total_chunks). */
int jn_claim_schedule_host(int total_chunks, int grid, int batch, int32_t* sizes /*HOST [6]*/,
                           int32_t* tickets /*HOST [7]*/, int32_t* chunks /*HOST [7]*/);

/* ------------------------------------------------------------------------------------------
 * K0  patch x bbox overlap tables.
 * ------------------------------------------------------------------------------------------ */
/* bboxes int64 [n, max_boxes, 4]; n_boxes int32 [n] or NULL (= all max_boxes rows are real);
 * rows/cols int32 [n] per-episode grid or NULL (= the scalars grid_rows/grid_cols);
 * out uint32 [n, words_per_item].  Replaces convert_bboxes_to_masks (general_env.py:360-379)
 * and bbox_positions (simple_env.py:270-321). */
int jn_patch_bitmaps(const int64_t* bboxes, const int32_t* n_boxes, int n, int max_boxes,
                     int patch_size, int grid_rows, int grid_cols, const int32_t* rows,
                     const int32_t* cols, int rule /*jn_overlap_rule*/, uint32_t* out,
                     int words_per_item, void* stream);
/* The same for boxes that are not whole pixels (float64 x1, y1, x2, y2; JN_RULE_AREA5 only): the dataset's
 * minimum-size resize scales the boxes (dataset.py:258-270) and NeedleSimpleEnv keeps them as python floats;
 * patch ranges floor(v / P), `oh * ow / P**2 > 0.05` and the centre floor((a + b) / 2) are evaluated in IEEE
 * doubles like the reference's float arithmetic (simple_env.py:13-18,270-321). */
int jn_patch_bitmaps_f64(const double* bboxes, const int32_t* n_boxes, int n, int max_boxes,
                         int patch_size, int grid_rows, int grid_cols, const int32_t* rows,
                         const int32_t* cols, int rule /*jn_overlap_rule*/, uint32_t* out,
                         int words_per_item, void* stream);
/* uint32 bitmaps -> one byte per patch, [n, rows, cols] (the reference's bool tensors). */
int jn_bitmap_unpack(const uint32_t* words, int n, int rows, int cols, uint8_t* out, void* stream);
/* Per-patch split of each box: local int64 [n, rows, cols, max_boxes, 4] (inclusive local
 * x1,y1,x2,y2, clamped to P-1) and present uint8 [n, rows, cols, max_boxes].  Replaces
 * parse_bboxes (general_env.py:381-504).  `status` gets bit 2 set when a box corner falls
 * outside the grid (the reference raises IndexError there). */
int jn_split_boxes(const int64_t* bboxes, int n, int max_boxes, int patch_size, int rows, int cols,
                   int64_t* local, uint8_t* present, int32_t* status, void* stream);
/* Per item the intersection of every raw box with patch (y, x) as float32
 * [n_items, max_boxes, 6] rows (0, x1, y1, x2, y2, 1) in local pixels, zero rows when empty.
 * episode = src_index ? src_index[i] : i (negative -> zero rows).  Replaces
 * NeedleSimpleEnv.local_bboxes (simple_env.py:231-268). */
int jn_local_boxes(const int64_t* bboxes, const int32_t* n_boxes, int max_boxes, int patch_size,
                   const int64_t* positions, const int32_t* src_index, int n_items, float* out,
                   void* stream);

/* float64 boxes: differences in double, rounded to float32 once (python floats into a FloatTensor). */
int jn_local_boxes_f64(const double* bboxes, const int32_t* n_boxes, int max_boxes, int patch_size,
                       const int64_t* positions, const int32_t* src_index, int n_items, float* out,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  batched env reset / step (a warp per 32 episodes: lane per episode for the scalar state, groups of
 * lanes over the bitmap words).
 * ------------------------------------------------------------------------------------------ */
/* Clears visited / steps / has_stopped and marks the start patch.  Replaces
 * init_env_variables + the tail of reset (general_env.py:117-142,164). */
int jn_env_reset(const int64_t* positions, uint32_t* visited, int64_t* steps, uint8_t* has_stopped,
                 int n, int rows, int cols, int32_t* status, void* stream);
/* One env step, in the reference's order (general_env.py:172-207): move + clamp, sticky STOP
 * flag, reward from the visited map BEFORE marking, mark, steps += 1, truncated, terminated.
 *   reward = fl32(fl32(fresh + fl32(-1/T)) + stop_eval)      (general_env.py:321-358)
 * `cost` is the host-rounded float32 of -1/max_ep_len.  pos_in and pos_out may alias.
 * `status` gets bit 1 set when an action code is outside [0, 8]. */
int jn_env_step(const int64_t* pos_in, const int64_t* actions, int64_t* pos_out, uint32_t* visited,
                const uint32_t* bbox, int64_t* steps, uint8_t* has_stopped, float* rewards,
                uint8_t* terminated, uint8_t* truncated, int n, int rows, int cols,
                int max_ep_len, float cost, int stop_enabled, int32_t* status, void* stream);
/* The `rewards` property evaluated on the current state, outside of a step (general_env.py:321-358):
 * float32 [n], same arithmetic as jn_env_step but with `visited` as it is now. */
int jn_env_rewards(const int64_t* positions, const uint32_t* visited, const uint32_t* bbox,
                   const uint8_t* has_stopped, int n, int rows, int cols, float cost, int stop_enabled,
                   float* rewards, void* stream);
/* prop_patches_found (general_env.py:308-315) as float32 [n]; `terminated` (general_env.py:
 * 235-246, uint8 [n]) is written too when non-NULL. */
int jn_env_props(const uint32_t* visited, const uint32_t* bbox, const uint8_t* has_stopped, int n,
                 int rows, int cols, int stop_enabled, float* prop_patches, uint8_t* terminated,
                 void* stream);

/* One call per env step: K2 followed by K1 without returning to the host language in between
 * (general_env.py:172-207 `step`, :144-170 `reset`).
 *
 * jn_env_step_gather launches the gather of the new glimpses -- tile (pos_out[i]) of image i ->
 * out + i * out_item_stride_bytes -- and the step kernel (as jn_env_step) as a programmatic-dependent-launch
 * pair.  The gather does not need the step kernel: it moves pos_in by the action itself (same move + clamp).
 * It is launched first and releases its dependents in its prologue, so the state update runs next to it and
 * a step costs one launch latency.  pos_in and pos_out must not alias.  `set` NULL or `out` NULL: state
 * update only.
 *
 * With a first-visit table (first_slot / host_src / history_src non-NULL; images in pinned HOST memory, crops
 * kept in a history buffer of `slots` slots per episode, `t` = the slot being written) the step kernel also
 * does the work of jn_visit_sources, the gather waits for it and reads only first visits from `set`, and a
 * second gather copies revisited patches out of `history_set` (the history registered as n * slots
 * one-patch images).  `host_tiles` (device uint64[1] or NULL) accumulates the number of first visits.
 *
 * jn_env_reset_gather: pos_out holds the start positions; clears the state (as jn_env_reset), initialises the
 * first-visit table when there is one (slot 0), gathers slot 0. */
typedef struct jn_env_step_args {
  /* episode state, borrowed device pointers */
  const int64_t* pos_in;  /* [n, 2] before the move (unused by reset) */
  const int64_t* actions; /* [n] */
  int64_t* pos_out;       /* [n, 2] after the move */
  uint32_t* visited;      /* [n, words] */
  const uint32_t* bbox;   /* [n, words] */
  int64_t* steps;         /* [n] */
  uint8_t* has_stopped;   /* [n] */
  float* rewards;         /* [n] */
  uint8_t* terminated;    /* [n] */
  uint8_t* truncated;     /* [n] */
  int32_t* first_slot;    /* [n, rows*cols] or NULL */
  int32_t* host_src;      /* [n] or NULL */
  int32_t* history_src;   /* [n] or NULL */
  unsigned long long* host_tiles; /* [1] or NULL */
  int32_t* status;        /* [1] or NULL: bit 1 bad position, bit 2 bad action */
  int32_t n, rows, cols, max_ep_len, stop_enabled, slots, t;
  float cost;             /* host-rounded float32 of -1 / max_ep_len */
  /* gather of the new glimpses (see jn_gather) */
  const int32_t* shifts;  /* [n_images, 2] or NULL */
  void* out;
  int64_t out_item_stride_bytes;
  uint32_t flags;
  int32_t engine;
} jn_env_step_args;

int jn_env_step_gather(const jn_images* set, const jn_images* history_set, const jn_env_step_args* args,
                       void* stream);
int jn_env_reset_gather(const jn_images* set, const jn_images* history_set, const jn_env_step_args* args,
                        void* stream);

/* Where this step's glimpse of every episode comes from when the images live in pinned HOST memory and the
 * crops are kept in a [n, slots, C, P, P] history buffer (NeedleGeneralEnv(history=True), the in-place form
 * of the trainer's concat, reinforce.py:175-179): first_slot int32 [n, rows*cols] (-1 = patch never seen)
 * records the history slot that first held each patch of each episode.  For the current positions
 * (int64 [n, 2]) and the slot `t` about to be written: a patch seen for the first time gets
 * host_src[i] = i, history_src[i] = JN_SRC_SKIP and is recorded; a revisited patch gets host_src[i] =
 * JN_SRC_SKIP and history_src[i] = i * slots + first_slot (the history as a set of one-patch images).
 * Two gathers then fill slot t; a patch crosses PCIe once per episode. */
int jn_visit_sources(const int64_t* positions, int32_t* first_slot, int n, int rows, int cols, int slots,
                     int t, int32_t* host_src, int32_t* history_src, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------
 * Glimpse pyramid: one level of NeedleGeneralEnv.init_glimps_images (general_env.py:84-115) --
 * TF.pad(level, [P]*4, "reflect") followed by TF.resize(.., [H, W], antialias=True) -- for float32 or uint8
 * images (`dtype`, jn_dtype).  src: n_images images of [C, H, W] pixels, src_image_stride_bytes apart; dst
 * likewise; tmp: n_images*C*H*W floats of scratch.  first / count / weights ([size, k] float32) per axis are the
 * antialiased bilinear filter taps of every output index over the PADDED axis (size + 2*pad -> size), computed
 * by the host exactly as ATen does (jolineedle_b200/pyramid.py:aa_weights).  Rows pass, then columns pass, every
 * output pixel the chain t = s0*w0, t = fma(s_j, w_j, t): bit-identical to torch's CPU kernel (AVX2 / AVX-512
 * builds).  uint8: torchvision casts to float32, resizes and casts back through torch.round (half to even); so
 * does this.
 * ------------------------------------------------------------------------------------------ */
int jn_resize_aa_reflect(const void* src, int64_t src_image_stride_bytes, float* tmp, void* dst,
                         int64_t dst_image_stride_bytes, int dtype /*jn_dtype*/, int n_images, int channels,
                         int height, int width, int pad, const int32_t* first_x, const int32_t* count_x,
                         const float* weights_x, int k_x, const int32_t* first_y, const int32_t* count_y,
                         const float* weights_y, int k_y, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3  segmented scans.
 * ------------------------------------------------------------------------------------------ */
/* Returns tail of a rollout (reinforce.py:186-202).  Inputs are step-major as the env
 * produces them: rewards float32 [T, n], terminated uint8 [T, n].  Outputs are episode-major
 * like the reference's stacked tensors: rewards_out [n, T], masks uint8 [n, T+1] (column 0
 * = 1, column t+1 = !terminated[t]), logit_masks uint8 [n, T], returns float32 [n, T] with
 * returns[:, t] = sum_{s >= t} rewards[:, s] * logit_masks[:, s], accumulated in float64
 * from the last step and rounded once per element (torch CPU cumsum semantics). */
int jn_returns(const float* rewards_tn, const uint8_t* terminated_tn, int T, int n,
               float* rewards_out, uint8_t* masks, uint8_t* logit_masks, float* returns,
               void* stream);
/* Same scan on episode-major inputs: rewards [n, T] (row stride in elements), logit_masks
 * uint8 [n, T]. */
int jn_returns_rows(const float* rewards, int64_t rewards_row_stride, const uint8_t* logit_masks,
                    int64_t masks_row_stride, int T, int n, float* returns, void* stream);

/* Supervised trajectories from supplied keypoints (simple_env.py:481-664).
 *
 * Episode e walks from start[e] through its segments seg_begin[e] .. seg_begin[e+1]-1.
 * Segment k is a straight-line walk (8-neighbour, diagonal first) to (seg_to_y, seg_to_x);
 * the recorded best action points at (seg_tgt_y, seg_tgt_x); seg_flags bit 0 marks the first
 * segment of a keypoint group (the "previous best action" overwrite, simple_env.py:548-552).
 * Whenever a best action would be STOP the next pre-drawn replacement move is consumed from
 * draws[draw_begin[e] ..] in the reference's draw order (simple_env.py:715-718).
 * Episodes longer than T keep their LAST T records (simple_env.py:573-584); shorter ones are
 * zero-padded with masks = 0.
 *
 * Outputs, all [n, T, ...]: positions int64 [.,2], current_actions / next_actions / labels
 * int64, masks float32, gather_src int32 (= e for recorded slots, -1 for padding; feed it
 * to jn_gather / jn_local_boxes together with `positions`), ep_len int32 [n] (untruncated).
 * labels = membership of the position in `area_bitmaps` (JN_RULE_AREA5 words, per-episode
 * grid cols from `cols`). */
int jn_traj_expand(const int32_t* start_yx /*[n,2]*/, const int32_t* seg_begin /*[n+1]*/,
                   const int32_t* seg_to_yx /*[S,2]*/, const int32_t* seg_tgt_yx /*[S,2]*/,
                   const uint8_t* seg_flags /*[S]*/, const int32_t* draw_begin /*[n+1]*/,
                   const uint8_t* draws, const uint32_t* area_bitmaps, int words_per_item,
                   const int32_t* cols /*[n]*/, int n, int T, int64_t* positions,
                   int64_t* current_actions, int64_t* next_actions, int64_t* labels, float* masks,
                   int32_t* gather_src, int32_t* ep_len, int32_t* status, void* stream);

/* Redirects gather queries to tiles that a previous gather already produced.  Trajectory records
 * (traj_positions int64 [n*T, 2], traj_src int32 [n*T] = episode or -1, as written by
 * jn_traj_expand) were gathered into a [n*T, C, P, P] buffer that is registered in the image set as a
 * slab of n*T one-patch images starting at image index `slab_base`.  Query d (episode query_src[d],
 * patch query_positions[d]) that matches a recorded slot of its episode becomes
 * (image slab_base + e*T + t, patch (0, 0)); other queries are passed through.  Used for the
 * detection patches of init_sample (simple_env.py:397-419), which mostly repeat trajectory glimpses:
 * with host-resident images those tiles then stay off PCIe. */
int jn_tile_lookup(const int64_t* traj_positions, const int32_t* traj_src, int T,
                   const int64_t* query_positions, const int32_t* query_src, int n_queries,
                   int slab_base, int64_t* out_positions, int32_t* out_src, void* stream);

/* First occurrences and repeats among the recorded slots of each episode (traj_positions / traj_src as
 * written by jn_traj_expand, n_slots = n * T).  A slot whose patch was already recorded at an earlier
 * slot of its episode gets first_src = JN_SRC_SKIP and repeat_src = index of that earlier slot; all
 * other slots keep traj_src in first_src and get repeat_src = JN_SRC_SKIP.  Gathering first_src out of
 * the images and then repeat_src out of the [n*T, C, P, P] buffer itself (as a set of one-patch images,
 * patch (0, 0)) yields the same buffer as one gather of traj_src (generate_sample's crops,
 * simple_env.py:560-572), but a revisited tile of a host-resident image crosses PCIe once. */
int jn_tile_dedupe(const int64_t* traj_positions, const int32_t* traj_src, int n_slots, int T,
                   int32_t* first_src, int32_t* repeat_src, void* stream);

/* ------------------------------------------------------------------------------------------
 * Host-side planner of supervised episodes (no GPU involved; all pointers are HOST memory).
 *
 * The host half of NeedleSimpleEnv.generate_sample (simple_env.py:378-441,481-629,666-718):
 * detection-patch pick, start position, greedy key-point order, random key points, detours and
 * the STOP replacement moves -- with the reference's random streams reproduced bit for bit
 * (numpy SeedSequence/PCG64/Generator, CPython random.choice on MT19937, CPython set iteration
 * order).  Its output is the input of jn_traj_expand / jn_gather / jn_local_boxes.
 *
 *   boxes int64 [n, max_boxes, 4] (x1,y1,x2,y2), n_boxes int32 [n], rows / cols int32 [n];
 *   seeds uint64 [n] + has_seed uint8 [n] (0 = draw OS entropy, like an unseeded numpy Generator);
 *   start_yx int32 [n, 2] or NULL (= drawn: y then x);
 *   mt_state uint32 [625]: the MT19937 words + index of python's global `random`
 *   (random.getstate()[1]); updated in place so the caller can random.setstate() afterwards.
 *
 * Binomial key points follow numpy (inversion up to n*p = 30, BTPE above).  JN_ERR_UNSUPPORTED for grids wider than
 * 4096 patches.
 * ------------------------------------------------------------------------------------------ */
typedef struct jn_plan jn_plan;
int jn_plan_create(jn_plan** out);
void jn_plan_destroy(jn_plan* plan);
const char* jn_plan_error(const jn_plan* plan);
int jn_plan_run(jn_plan* plan, int n, const int64_t* boxes, const int32_t* n_boxes, int max_boxes,
                const int32_t* rows, const int32_t* cols, int patch_size, const uint64_t* seeds,
                const uint8_t* has_seed, int min_keypoints, int max_keypoints, int binomial,
                const int32_t* start_yx, uint32_t* mt_state);
/* The same on a thread of the library's own: jn_plan_start returns at once, jn_plan_wait joins and returns
 * jn_plan_run's status.  The caller keeps every input array alive and untouched in between and calls nothing
 * else on the plan; a host language with a global interpreter lock can so overlap the planning of a batch with
 * its own work on the same batch (building the image set). */
int jn_plan_start(jn_plan* plan, int n, const int64_t* boxes, const int32_t* n_boxes, int max_boxes,
                  const int32_t* rows, const int32_t* cols, int patch_size, const uint64_t* seeds,
                  const uint8_t* has_seed, int min_keypoints, int max_keypoints, int binomial,
                  const int32_t* start_yx, uint32_t* mt_state);
int jn_plan_wait(jn_plan* plan);
int jn_plan_sizes(const jn_plan* plan, int* n_segments, int* n_draws, int* n_det);
/* start [n,2], seg_begin [n+1], seg_to [S,2], seg_tgt [S,2], draw_begin [n+1], det_begin [n+1],
 * det_yx [D,2] (int32); seg_flags [S], draws [Q] (uint8).  NULL pointers are skipped. */
int jn_plan_export(const jn_plan* plan, int32_t* start, int32_t* seg_begin, int32_t* seg_to,
                   int32_t* seg_tgt, int32_t* draw_begin, int32_t* det_begin, int32_t* det_yx,
                   uint8_t* seg_flags, uint8_t* draws);

#ifdef __cplusplus
}
#endif
#endif /* JOLINEEDLE_B200_H */
