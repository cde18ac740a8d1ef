/*
 * muvo_b200.h -- C ABI of libmuvo_b200.so: B200 (sm_100a) kernels for MUVO's
 * geometric sensor-to-grid hot path.
 *
 * The reference (fzi-forschungszentrum-informatik/muvo) is pure Python and has no
 * FFI layer; each entry point below names the reference function it replaces
 * (file:line, relative to the MUVO checkout).  The Python wrappers in
 * muvo_b200/ keep the reference call signatures and bind these symbols with
 * ctypes; INTEGRATION.md shows the binding a MUVO maintainer would add.
 *
 * Conventions
 *  - Pointers are DEVICE pointers unless the name ends in _h (host).
 *  - The caller owns every buffer (inputs, outputs, workspace); nothing is
 *    allocated or freed inside the library and no global mutable state is kept
 *    (except the debug tuning knobs at the end of this header).
 *  - Every call enqueues work on `stream` (a cudaStream_t passed as void*) and
 *    returns without synchronising.  Calls are re-entrant; concurrent callers
 *    must use distinct workspaces.
 *  - Return value: 0 = OK, <0 = MUVO_E_* below, >0 = cudaError_t of a launch.
 *    muvo_strerror() maps both to text.
 *  - Workspaces are "self-cleaning": muvo_ws_reset() must be enqueued once
 *    after allocation, after any failed call, and whenever (n_frames, grid
 *    size, H*W) differ from the previous call on that workspace (the table
 *    layout depends on them; the number of points may vary freely).  Every
 *    successful call leaves the workspace ready for the next one.
 *  - Frames are ragged: point p of frame f lives at rows
 *    [frame_offsets[f], frame_offsets[f+1]) of the packed point arrays.
 */
#ifndef MUVO_B200_H
#define MUVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MUVO_B200_ABI_VERSION 5

#if defined(__GNUC__)
#define MUVO_API __attribute__((visibility("default")))
#else
#define MUVO_API
#endif

enum {
  MUVO_OK = 0,
  MUVO_E_NULL = -1,      /* required pointer is NULL */
  MUVO_E_ARG = -2,       /* invalid scalar argument / unsupported dtype */
  MUVO_E_SHAPE = -3,     /* shape overflow (e.g. > 2^31 voxels per frame) */
  MUVO_E_WORKSPACE = -4, /* workspace too small */
  MUVO_E_ALIGN = -5      /* pointer not aligned as documented */
};

enum { MUVO_F32 = 0, MUVO_F64 = 1, MUVO_F16 = 2, MUVO_BF16 = 3 };                 /* floating dtypes */
enum { MUVO_I64 = 0, MUVO_I32 = 1, MUVO_U8 = 2, MUVO_I16 = 3 };                   /* integer dtypes (pred) */
enum { MUVO_RANGE_LAYOUT_HWC = 0, /* depth[F,H,W] f32, xyz[F,H,W,3] f32, sem[F,H,W] u8: do_range_projection's returns */
       MUVO_RANGE_LAYOUT_XYZD = 1 /* xyz := [F,4,H,W] f32 planes x,y,z,depth (dataset.py:301-303); depth may be NULL   */ };

/* Occupancy-grid description.  Replaces the (voxel_resolution, voxel_size, offset)
 * arguments of voxel_filter(), data/data_preprocessing.py:172-178.  The host
 * computes offset/upper in float64 exactly as numpy does there:
 *   offset = user_offset + res*size/2 ; upper = size*res.                          */
typedef struct MuvoGrid {
  double res;
  double offset[3];
  double upper[3];
  int32_t size[3];      /* Dx, Dy, Dz */
  int32_t roadline_id;  /* label that overrides the nearest-point label (6, :198); <0 disables */
} MuvoGrid;

/* Range-image description.  Replaces PointCloud.__init__, muvo/utils/geometry_utils.py:167-173. */
typedef struct MuvoRangeCfg {
  int32_t H, W;
  double fov_down_abs;  /* abs(fov_down) in rad          (:169,:190) */
  double fov;           /* fov_up - fov_down in rad      (:170)      */
  double lidar_pos[3];  /* (:173,:178)                               */
} MuvoRangeCfg;

/* LiDAR-side prep in front of the range projection, muvo/data/dataset.py:275-290:
 * convert_coor_lidar (data/data_preprocessing.py:119-122: p = float32(float64(p) + add), then y = -y), the ego-box drop
 * (:286-290: points with box_lo < p < box_hi on all three axes are removed before the projection) and the LABEL_MAP
 * remap of the point semantics (:281-283, applied to the winner's tag when the image is written).                       */
typedef struct MuvoLidarPrep {
  double add[3];            /* cfg.POINTS.LIDAR_POSITION                                   */
  double box_lo[3];         /* (-x/2, -y/2, 0) of EGO_VEHICLE_DIMENSION                     */
  double box_hi[3];         /* ( x/2,  y/2, z)                                              */
  int32_t use_ego_box;      /* 0: keep every point                                          */
  int32_t reserved;
  const uint8_t* remap256;  /* DEVICE pointer to a 256-entry table, or NULL = tags as they are */
} MuvoLidarPrep;

/* diag[] slots written by the point kernels (int64 each) */
enum { MUVO_DIAG_DROPPED_NONFINITE = 0, /* points with NaN/Inf coordinates or at the sensor origin (reference: IndexError) */
       MUVO_DIAG_NEAR_EDGE_W = 1,       /* |frac(proj_w)| within 1e-9 of a column edge (incl. exactly on it)               */
       MUVO_DIAG_NEAR_EDGE_H = 2,       /* same for rows                                                                   */
       MUVO_DIAG_IN_GRID = 3,           /* points inside the occupancy grid                                                */
       MUVO_DIAG_COUNT = 8 };

MUVO_API int muvo_abi_version(void);
MUVO_API const char* muvo_strerror(int code);

/* ---- per-kernel timing (bench / profiling only) ------------------------------------
 * Between begin and end, every kernel launched by this host thread through the library is followed
 * by a CUDA event on its stream.  end() synchronises the stream and returns, per launch in order,
 * the elapsed milliseconds since the previous mark and the kernel's name (static strings).
 * Thread-local; at most 63 launches are recorded.                                       */
MUVO_API int muvo_profile_begin(void* stream);
MUVO_API int muvo_profile_end(void* stream, int32_t capacity, float* ms_out_h, const char** names_out_h,
                              int32_t* n_out_h);

/* ---- host staging ---------------------------------------------------------------
 * memcpy of `bytes` split over up to n_threads host threads (0 = one per core, capped): fills pinned staging buffers
 * from the application's pageable arrays at several times the single-thread rate.  Blocking; no CUDA call inside. */
MUVO_API int muvo_host_copy(void* dst_h, const void* src_h, size_t bytes, int32_t n_threads);

/* ---- workspace ------------------------------------------------------------------ */
/* Bytes needed by the point kernels for `n_frames` frames / `n_points_total` points.
 * grid_h or range_h may be NULL when that stage is not used.                        */
MUVO_API int muvo_points_workspace_bytes(int64_t n_points_total, int32_t n_frames, const MuvoGrid* grid_h,
                                const MuvoRangeCfg* range_h, size_t* bytes_out_h);
/* Put a freshly allocated (or dirty) workspace into the clean state. */
MUVO_API int muvo_ws_reset(void* ws, size_t ws_bytes, void* stream);

/* ---- N1: camera + LiDAR cloud in front of (a) -------------------------------------
 * Replaces merge_pcd(), data/data_preprocessing.py:125-139 (depth decode :72-77, depth2pcd :87-106, convert_coor_img
 * :109-119, convert_coor_lidar :121-123, ego-box mask :133-138), bit-exact in float64, order preserving.
 *   img_bgra  [H,W,4] uint8 as cv2.imread(file, -1) returns the encoded depth + semantic image
 *   focal = W / (2 tan(fov pi / 360)) (the host evaluates it as numpy does), range = 100 (:87)
 *   camera_pos_h [3] (forward, right, up; float32 values, :111), lidar_pos_h [3], ego_box_h [6] = lo xyz, hi xyz or NULL
 *   lidar_xyz [N,3] float32 in the LiDAR frame, lidar_sem [N] uint8
 *   xyz_out [H*W + N, 3] float64, sem_out [H*W + N] uint8 (first *n_out rows valid), n_out [1] int64 (device)        */
MUVO_API int muvo_merge_pcd_workspace_bytes(int32_t H, int32_t W, int64_t n_lidar, size_t* bytes_out_h);
MUVO_API int muvo_merge_pcd(const uint8_t* img_bgra, int32_t H, int32_t W, double focal, double range,
                   const double* camera_pos_h, const float* lidar_xyz, const uint8_t* lidar_sem, int64_t n_lidar,
                   const double* lidar_pos_h, const double* ego_box_h, double* xyz_out, uint8_t* sem_out,
                   int64_t* n_out, void* ws, size_t ws_bytes, void* stream);
/* The same for frame `frame` of a batch: the merged cloud is written at row row_offsets[frame] of xyz_out / sem_out (device
 * int64 array, row_offsets[0] set by the caller, normally 0) and row_offsets[frame + 1] = row_offsets[frame] + n is written by
 * the call, so that N frames are packed back to back by N stream-ordered calls without a host synchronisation; the array is
 * then the frame_offsets of muvo_voxelize.  Rows at or beyond capacity_rows are not written (check row_offsets afterwards). */
MUVO_API int muvo_merge_pcd_at(const uint8_t* img_bgra, int32_t H, int32_t W, double focal, double range, const double* camera_pos_h,
                      const float* lidar_xyz, const uint8_t* lidar_sem, int64_t n_lidar, const double* lidar_pos_h,
                      const double* ego_box_h, double* xyz_out, uint8_t* sem_out, int64_t capacity_rows, int64_t* row_offsets,
                      int32_t frame, void* ws, size_t ws_bytes, void* stream);

/* ---- N1: label pyramids behind (a)/(b) --------------------------------------------
 * Replaces the LIDAR_RE / LIDAR_SEG / VOXEL_SEG blocks of PreProcess.forward, muvo/models/preprocess.py:151-186:
 * rv1 = range_xyzd / scale, rv2 / rv4 = nearest-neighbour halvings (H/2 x W/2, H/4 x W/4), seg2 / seg4 the same for
 * range_sem, vox2 / vox4 for the voxel grid (X/2 x Y/2 x Z/2, /4).  Index rule = PyTorch `nearest`.  Either input family
 * may be NULL (then its outputs are ignored).                                             */
MUVO_API int muvo_label_pyramids(const float* range_xyzd, const uint8_t* range_sem, const uint8_t* voxel, int32_t F, int32_t H,
                        int32_t W, int32_t X, int32_t Y, int32_t Z, float scale, float* rv1, float* rv2, float* rv4,
                        uint8_t* seg2, uint8_t* seg4, uint8_t* vox2, uint8_t* vox4, void* stream);

/* Saved sparse voxels -> dense training grids: muvo/data/dataset.py:317-327 for F files at once.
 *   rows [n_rows,4] uint16 (x, y, z, label) as data/generate_voxels.py:72-73 saves them, files back to back,
 *   row_offsets [F+1] int64; label 255 -> 0 (:322), remap256[label] (:323, NULL = identity); the LAST row of a voxel wins
 *   like numpy's fancy assignment (:325); rows outside the grid are skipped and counted in n_bad [1] int64 (numpy raises
 *   IndexError there; may be NULL).  dense_out [F,Dx,Dy,Dz] uint8 fully written, 4-byte aligned; scratch_rows [n_rows] uint8. */
MUVO_API int muvo_densify_sparse(const uint16_t* rows, const int64_t* row_offsets, int32_t n_frames, int64_t n_rows, int32_t dx,
                        int32_t dy, int32_t dz, const uint8_t* remap256, uint8_t* dense_out, uint8_t* scratch_rows,
                        int64_t* n_bad, void* stream);

/* ---- (a) voxelisation ------------------------------------------------------------
 * Replaces voxel_filter(), data/data_preprocessing.py:172-228, batched over frames, and
 * (dense_out) the densify step of muvo/data/dataset.py:317-327.
 *   xyz        [P,3]  float32 or float64 (xyz_dtype), ego frame
 *   sem        [P]    uint8
 *   remap256   [256]  uint8 label remap applied to dense_out only, or NULL
 *   dense_out  [F,Dx,Dy,Dz] uint8 (fully written, 0 = empty), or NULL
 *   sparse_out [P,4]  uint16 rows (x,y,z,label), ordered by x + y*Dx + z*Dx*Dy (:184,:187), or NULL; frame f's
 *                     n_occ[f] rows start at row frame_offsets[f], or -- when sparse_start_out is given -- at row
 *                     sparse_start_out[f]: the frames' lists back to back (a read-back then moves only the rows used)
 *   n_occ_out  [F]    int64 occupied voxels per frame, or NULL
 *   sparse_start_out [F+1] int64 exclusive prefix of n_occ (needs sparse_out and n_occ_out), or NULL
 *   diag       [MUVO_DIAG_COUNT] int64, accumulated into (caller zeroes), or NULL        */
MUVO_API int muvo_voxelize(const void* xyz, int32_t xyz_dtype, const uint8_t* sem, const int64_t* frame_offsets,
                  int32_t n_frames, int64_t n_points_total, const MuvoGrid* grid_h, const uint8_t* remap256,
                  uint8_t* dense_out, uint16_t* sparse_out, int64_t* n_occ_out, int64_t* sparse_start_out,
                  int64_t* diag, void* ws, size_t ws_bytes, void* stream);

/* ---- (b) range-view projection ---------------------------------------------------
 * Replaces PointCloud.do_range_projection(), muvo/utils/geometry_utils.py:175-220.
 *   xyz [P,3] float32 ego frame; outputs per `layout` above (all pixels written:
 *   empty = depth -1, xyz 0, sem 0).                                                */
MUVO_API int muvo_range_project(const float* xyz, const uint8_t* sem, const int64_t* frame_offsets, int32_t n_frames,
                       int64_t n_points_total, const MuvoRangeCfg* cfg_h, int32_t layout, float* depth_out,
                       float* xyz_out, uint8_t* sem_out, int64_t* diag, void* ws, size_t ws_bytes, void* stream);

/* (b) straight from the raw semantic-LiDAR sweep: what muvo/data/dataset.py:275-300 does per sample --
 * convert_coor_lidar, LABEL_MAP remap, ego-box drop, do_range_projection -- as ONE pass of the same kernels (the prep is applied
 * where a point is loaded; no intermediate cloud is written).  xyz_raw [P,3] float32 in the LiDAR frame, tag [P] uint8 raw
 * CARLA tags; outputs as muvo_range_project (range_xyz holds the converted ego-frame points, range_sem the remapped tags). */
MUVO_API int muvo_range_project_lidar(const float* xyz_raw, const uint8_t* tag, const int64_t* frame_offsets, int32_t n_frames,
                       int64_t n_points_total, const MuvoRangeCfg* cfg_h, const MuvoLidarPrep* prep_h, int32_t layout,
                       float* depth_out, float* xyz_out, uint8_t* sem_out, int64_t* diag, void* ws, size_t ws_bytes,
                       void* stream);

/* ---- (a)+(b) fused: one read of the point stream feeds both stages ---------------- */
MUVO_API int muvo_points_fused(const float* xyz, const uint8_t* sem, const int64_t* frame_offsets, int32_t n_frames,
                      int64_t n_points_total, const MuvoGrid* grid_h, const uint8_t* remap256,
                      const MuvoRangeCfg* cfg_h, int32_t layout, uint8_t* dense_out, uint16_t* sparse_out,
                      int64_t* n_occ_out, int64_t* sparse_start_out, float* depth_out, float* xyz_out, uint8_t* sem_out,
                      int64_t* diag, void* ws, size_t ws_bytes, void* stream);

/* ---- (c) lift-splat BEV pooling ---------------------------------------------------
 * Replaces FrustumPooling.voxel_pooling + QuickCumsum (muvo/models/frustum_pooling.py:34-60,
 * 131-187; twin: VoxelsSumming, muvo/layers/layers.py:326-357).
 *   x        lifted features viewed as [B, n_pts, C] with element strides
 *            (x_stride_b, x_stride_p, x_stride_c); n_pts = N*D*H*W frustum points per frame
 *   cell     [B, n_pts] int32 flat BEV cell ((z*ny + y)*nx + x) or -1 = dropped
 *            (masked out or out of bounds, :153-163)
 *   out      [B, C*nz, ny, nx] float32, fully written (empty cells = 0, :180-185)
 * The sum over a cell's points is taken in ascending point order (deterministic).   */
MUVO_API int muvo_bev_pool_workspace_bytes(int32_t B, int64_t n_pts, int32_t n_cells, size_t* bytes_out_h);
/* Largest n_cells the forward kernels accept (one histogram per warp in shared memory: 12800 cells, e.g. up to a 113 x 113
 * BEV with nz = 1; MUVO's muvo.yml uses 48 x 48 = 2304).  Larger grids return MUVO_E_SHAPE.                              */
MUVO_API int muvo_bev_pool_max_cells(void);
/* cell_out[i] = mask[i] ? cell0[i] : -1 for i < n: the top-k depth mask of mile.py:512-514 (frustum_pooling.py:153-156)
 * applied to mask-independent cell ids, which the host caches per (intrinsics, extrinsics, frustum shape).              */
MUVO_API int muvo_bev_fold_mask(const int32_t* cell0, const uint8_t* mask, int64_t n, int32_t* cell_out, void* stream);
MUVO_API int muvo_bev_pool_fwd(const void* x, int32_t x_dtype, int64_t x_stride_b, int64_t x_stride_p, int64_t x_stride_c,
                      const int32_t* cell, int32_t B, int64_t n_pts, int32_t C, int32_t n_cells, float* out,
                      void* ws, size_t ws_bytes, void* stream);
/* FrustumPooling.forward as the module calls it (frustum_pooling.py:189-209 with the mask of mile.py:512-514): cell0 = the
 * cached mask-independent cell ids, mask = the top-k depth mask as uint8 [B, n_pts] (nullptr = no mask), cell_out [B, n_pts]
 * receives cell0 with the mask folded in (what muvo_bev_pool_bwd needs).  When x is (B, C, D, H, W) memory (x_stride_p == 1,
 * 16-byte aligned rows) the tensor is streamed through shared memory with TMA bulk copies instead of gathered
 * (bev_stream.cu); otherwise this is muvo_bev_fold_mask + muvo_bev_pool_fwd.  Same output contract as muvo_bev_pool_fwd. */
MUVO_API int muvo_bev_pool_fwd_masked(const void* x, int32_t x_dtype, int64_t x_stride_b, int64_t x_stride_p, int64_t x_stride_c,
                      const int32_t* cell0, const uint8_t* mask, int32_t* cell_out, const void* plan, int32_t B, int64_t n_pts,
                      int32_t C, int32_t n_cells, float* out, void* ws, size_t ws_bytes, void* stream);
/* The mask-independent half of the streamed forward, cached by the host next to cell0 (per intrinsics / extrinsics / frustum
 * shape, frustum_pooling.py:111-163): per chunk of 2048 frustum points the kept points' (cell, position) keys in cell-sorted
 * order.  `plan` (muvo_bev_plan_bytes bytes, 256-byte aligned, caller-owned) may be passed to muvo_bev_pool_fwd_masked
 * together with the SAME cell0: per call only a stable compaction by the mask remains (no sort).  plan = nullptr: the lists
 * are sorted per call.                                                                                                     */
MUVO_API int muvo_bev_plan_bytes(int32_t B, int64_t n_pts, size_t* bytes_out_h);
MUVO_API int muvo_bev_plan_build(const int32_t* cell0, int32_t B, int64_t n_pts, int32_t n_cells, void* plan, size_t plan_bytes,
                      void* stream);
/* grad_x[b,p,c] = cell[b,p] >= 0 ? grad_out[b,c,cell[b,p]] : 0  (frustum_pooling.py:52-60 + index bwd);
 * grad_x is written with the given element strides (every element written).         */
MUVO_API int muvo_bev_pool_bwd(const float* grad_out, const int32_t* cell, int32_t B, int64_t n_pts, int32_t C,
                      int32_t n_cells, void* grad_x, int32_t gx_dtype, int64_t gx_stride_b, int64_t gx_stride_p,
                      int64_t gx_stride_c, void* stream);
/* 1 when muvo_bev_pool_fwd / _fwd_masked stream a tensor of this layout (bev_stream.cu), 0 when they gather. */
MUVO_API int muvo_bev_pool_is_streamed(int32_t elem_bytes, const void* x, int64_t x_stride_b, int64_t x_stride_p, int64_t x_stride_c,
                      int32_t B, int64_t n_pts, int32_t C, int32_t n_cells);
/* muvo_bev_pool_bwd for a gradient tensor in (B, C, D, H, W) memory: 8-channel tiles are assembled in shared memory (zeros +
 * the kept points, found through the chunk lists the STREAMED forward left in its workspace) and written with TMA bulk stores.
 * fwd_ws = the workspace of the matching forward call, untouched since, and only if muvo_bev_pool_is_streamed() held for that
 * forward's x; pass NULL (or an ineligible grad_x layout) and this is muvo_bev_pool_bwd.                                 */
MUVO_API int muvo_bev_pool_bwd_streamed(const float* grad_out, const int32_t* cell, int32_t B, int64_t n_pts, int32_t C,
                      int32_t n_cells, void* grad_x, int32_t gx_dtype, int64_t gx_stride_b, int64_t gx_stride_p,
                      int64_t gx_stride_c, const void* fwd_ws, size_t fwd_ws_bytes, void* stream);

/* Fused lift-splat (opt-in replacement of the two steps at muvo/models/mile.py:517-523):
 * out[b,c,cell] = sum over the kept frustum points p = (d, hw) of the cell, ascending p, of depth[b,d,hw] * feat[b,hw,c]
 * -- the lifted tensor (236 MB per frame at muvo.yml shapes) is never written.
 *   feat_cl [B, HW, C] float32 (channels-last image features), depth [B, D, HW] float32, cell [B, D*HW] int32 (-1 = dropped)
 *   out [B, C, n_cells] float32 fully written; workspace = muvo_bev_pool_workspace_bytes(B, D*HW, n_cells).
 * Backward: gout_cl [B, n_cells, C]; grad_depth [B, D, HW] (0 at dropped points) and grad_feat_cl [B, HW, C], both fully
 * written, deterministic (no atomics).                                                     */
MUVO_API int muvo_lift_splat_fwd(const float* feat_cl, const float* depth, const int32_t* cell, int32_t B, int32_t D,
                        int32_t HW, int32_t C, int32_t n_cells, float* out, void* ws, size_t ws_bytes, void* stream);
/* The same forward with the mask-independent cell sort taken from a cached plan (one per camera rig, built from the cell ids
 * WITHOUT the mask folded in): plan = muvo_lift_splat_plan_bytes(B, D*HW, n_cells) bytes, 256-byte aligned, filled by
 * muvo_lift_splat_plan_build (ws = muvo_bev_pool_workspace_bytes).  mask: uint8 [B, D*HW] (0 = dropped) or NULL; a call then
 * only filters the plan by the mask (two small kernels) instead of sorting.  Same output bits as muvo_lift_splat_fwd on the
 * folded cell ids.                                                                                                          */
MUVO_API int muvo_lift_splat_plan_bytes(int32_t B, int64_t n_pts, int32_t n_cells, size_t* bytes_out_h);
MUVO_API int muvo_lift_splat_plan_build(const int32_t* cell0, int32_t B, int64_t n_pts, int32_t n_cells, void* plan, size_t plan_bytes,
                      void* ws, size_t ws_bytes, void* stream);
MUVO_API int muvo_lift_splat_fwd_planned(const float* feat_cl, const float* depth, const void* plan, size_t plan_bytes, const uint8_t* mask,
                      int32_t B, int32_t D, int32_t HW, int32_t C, int32_t n_cells, float* out, void* ws, size_t ws_bytes, void* stream);
MUVO_API int muvo_lift_splat_bwd(const float* gout_cl, const float* feat_cl, const float* depth, const int32_t* cell,
                        int32_t B, int32_t D, int32_t HW, int32_t C, int32_t n_cells, float* grad_depth,
                        float* grad_feat_cl, void* stream);

/* Sorted-rank segment sum: QuickCumsum.forward / cumsum_trick (frustum_pooling.py:23-42).
 *   x [n,C] float32 row-major, ranks [n] int64 non-decreasing.
 *   seg_id_out [n] int32 (segment index of each row; scratch + used by the backward),
 *   n_seg_out  [1] int32, x_seg_out [>=n_seg, C] (caller sizes it for n rows),
 *   last_row_out [>=n_seg] int64 = index of the LAST row of each segment (geom_feats[kept], :40). */
MUVO_API int muvo_segment_sum_workspace_bytes(int64_t n, size_t* bytes_out_h);
MUVO_API int muvo_segment_sum_fwd(const float* x, const int64_t* ranks, int64_t n, int32_t C, int32_t* seg_id_out,
                         int32_t* n_seg_out, float* x_seg_out, int64_t* last_row_out, void* ws, size_t ws_bytes,
                         void* stream);
/* grad_x[i,:] = grad_seg[seg_id[i],:]  (QuickCumsum.backward, :52-60) */
MUVO_API int muvo_segment_sum_bwd(const float* grad_seg, const int32_t* seg_id, int64_t n, int32_t C, float* grad_x,
                         void* stream);

/* ---- (d) SSC / IoU counts ---------------------------------------------------------
 * Replaces SSCMetrics.get_score_completion + get_score_semantic_and_completion,
 * muvo/metrics.py:143-216, summed over all frames.
 *   pred      [n] integer labels (pred_dtype), target [n] uint8
 *   nonempty  [n] uint8/bool or NULL : restricts completion AND per-class counts
 *   nonsurface[n] uint8/bool or NULL : restricts completion only (add_batch, :79-95)
 *   ignore255 != 0 : additionally drop target==255 voxels (add_batch's y_true != 255)
 *   counts_out[3+3C] int64, ACCUMULATED into (caller zeroes):
 *              completion tp,fp,fn ; tp[C] ; fp[C] ; fn[C]                            */
MUVO_API int muvo_ssc_counts(const void* pred, int32_t pred_dtype, const uint8_t* target, const uint8_t* nonempty,
                    const uint8_t* nonsurface, int32_t ignore255, int64_t n_voxels, int32_t n_classes,
                    int64_t* counts_out, void* stream);
/* Fused argmax + counts over logits [F, C, S] (S voxels per frame), trainer.py:483-490:
 * never materialises the int64 prediction.  target [F,S] uint8.                      */
MUVO_API int muvo_ssc_counts_from_logits(const void* logits, int32_t logits_dtype, const uint8_t* target, int32_t n_frames,
                                int32_t n_classes, int64_t voxels_per_frame, int32_t ignore255,
                                int64_t* counts_out, void* stream);

/* ---- next row N4: SemScalLoss / GeoScalLoss reductions -------------------------------------
 * Replaces the softmax + per-class masked reductions of SemScalLoss.forward / GeoScalLoss.forward,
 * muvo/losses.py:199-251 and :259-287 (called from trainer.py:375-382).  Both losses are functions of
 *   sums_out[3C+1] float64 = sum_p[C] (sum of softmax prob. of class i over valid voxels),
 *                            nom[C]   (same, restricted to target == i), cnt[C] (#valid voxels with target == i), n_valid
 * where valid = target != ignore_index (pass -1 for "no ignore").  logits [F, C, S] f32/f16/bf16 contiguous (softmax in
 * fp32 as under autocast), target [F, S] uint8, n_classes <= 32.  sums_out is OVERWRITTEN; the result is deterministic.
 * losses_out[2 + 4C] float64 (or NULL) additionally receives SemScalLoss, GeoScalLoss and their derivatives
 * d SemScal / d(sum_p[C], nom[C]), d GeoScal / d(sum_p[C], nom[C]), evaluated with the reference's conditions
 * (class skipped when absent, term skipped when outside [0,1], log clamped at -100; NaN where the reference raises).
 * muvo_scal_sums_bwd writes grad_logits [F, C, S] (logits' dtype, fully written) = g_sem * d SemScal / d logits +
 * g_geo * d GeoScal / d logits, from dloss = losses_out + 2 (the [2, 2C] derivative block of the forward call) and the two
 * upstream gradients g_sem, g_geo (device float32 scalars; NULL = 0).                                                    */
MUVO_API int muvo_scal_workspace_bytes(int32_t n_classes, size_t* bytes_out_h);
MUVO_API int muvo_scal_sums_fwd(const void* logits, int32_t logits_dtype, const uint8_t* target, int32_t n_frames,
                                int32_t n_classes, int64_t voxels_per_frame, int32_t ignore_index, double* sums_out,
                                double* losses_out, void* ws, size_t ws_bytes, void* stream);
MUVO_API int muvo_scal_sums_bwd(const void* logits, int32_t logits_dtype, const uint8_t* target, int32_t n_frames,
                                int32_t n_classes, int64_t voxels_per_frame, int32_t ignore_index, const double* dloss,
                                const float* g_sem, const float* g_geo, void* grad_logits, void* stream);

/* ---- next row N4: the torch_scatter reductions of the PointPillar encoder ---------------------
 * Replaces scatter_mean(xyz, inverse_indices, dim=0) and scatter_max(feat, inverse_indices, dim=0)
 * (muvo/models/common.py:731 and :703; torch_scatter is a third-party dependency of the reference, not vendored).
 *   src [n_src, n_feat] float32 row-major, index [n_src] int64 or int32 (index_dtype), values in [0, n_out)
 *   out [n_out, n_feat] float32, fully written: mean / max over the rows with index == m, 0 for rows that receive none
 *   scatter_mean: count_out [n_out] int32 (rows per output, kept for the backward); deterministic: 64-bit fixed-point sums
 *                 (integer atomics), rounded to float32 once; ws = muvo_pillar_workspace_bytes(n_out, n_feat), 8-byte aligned
 *   scatter_max : arg_out [n_out, n_feat] int64 or NULL: the LOWEST source row attaining the maximum, n_src for empty
 *                 rows; deterministic; ws = muvo_pillar_workspace_bytes(n_out, n_feat), 8-byte aligned
 *   bad_index_flag [1] int32 device: set to 1 when an index is out of range (that row is skipped); caller zeroes it
 *   *_bwd: grad_src [n_src, n_feat] fully written from grad_out [n_out, n_feat] (0 for rows whose index is out of range). */
MUVO_API int muvo_pillar_workspace_bytes(int64_t n_out, int32_t n_feat, size_t* bytes_out_h);
MUVO_API int muvo_pillar_scatter_mean(const float* src, const void* index, int32_t index_dtype, int64_t n_src, int32_t n_feat,
                                      int64_t n_out, float* out, int32_t* count_out, void* ws, size_t ws_bytes,
                                      int32_t* bad_index_flag, void* stream);
MUVO_API int muvo_pillar_scatter_mean_bwd(const float* grad_out, const void* index, int32_t index_dtype, const int32_t* count,
                                          int64_t n_src, int32_t n_feat, int64_t n_out, float* grad_src, void* stream);
MUVO_API int muvo_pillar_scatter_max(const float* src, const void* index, int32_t index_dtype, int64_t n_src, int32_t n_feat,
                                     int64_t n_out, float* out, int64_t* arg_out, void* ws, size_t ws_bytes,
                                     int32_t* bad_index_flag, void* stream);
MUVO_API int muvo_pillar_scatter_max_bwd(const float* grad_out, const void* index, int32_t index_dtype, const int64_t* arg,
                                         int64_t n_src, int32_t n_feat, int64_t n_out, float* grad_src, void* stream);

/* ---- test / tuning hooks (not part of the drop-in surface) ------------------------------
 * muvo_debug_pixel_check: runs the f32 pixel path of the range projection next to the float64 formula of
 * geometry_utils.py:180-200 on n_points float32 ego-frame points and ACCUMULATES into counts_out[4] (device, int64):
 * [0] points decided by the f32 path, [1] points deferred to the float64 formula (near a bin edge),
 * [2] f32-decided points whose pixel differs from the float64 one (must stay 0), [3] dropped (non-finite / at the sensor).
 * muvo_debug_set_tuning: process-wide launch knobs for benchmarking sweeps; key 0 = CTAs per SM of the persistent
 * point pass (0 = as many as fit), key 1 bit 0 = disable the neighbour filter in front of the voxel atomicMax,
 * key 2 = 1 forces the per-point gather kernel of the BEV pool forward (default: streamed rows); key 3 = 1 forces the
 * multi-launch point path where the single-launch dataflow kernel would be eligible.                                   */
MUVO_API int muvo_debug_pixel_check(const float* xyz, int64_t n_points, const MuvoRangeCfg* cfg_h, int64_t* counts_out,
                                    void* stream);
MUVO_API int muvo_debug_set_tuning(int32_t key, int32_t value);
/* muvo_debug_mega_stats: synchronises the device and copies the per-CTA cycle counters of the dataflow point kernel
 * (points_mega.cu: 12 uint64 per CTA -- waiting / P / ER / ED / queue / producer waits / unit counts / total) of the first
 * n_ctas CTAs into out_h (host, may be NULL); reset != 0 zeroes them afterwards.  Counters accumulate over calls.        */
MUVO_API int muvo_debug_mega_stats(unsigned long long* out_h, int32_t n_ctas, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* MUVO_B200_H */
