// Internal (C++ linkage) forms of the kernel families that td_chain_* composes.  They differ from the
// C-ABI entry points of include/treedet.h only in what a fully device-driven, graph-capturable chain
// needs: live counts, per-image parameters and ring indirections read from DEVICE memory, so that one
// captured launch sequence serves every image of a tiling.
#pragma once
#include "common.cuh"

// per-image parameters of the chain, uploaded once per image (pinned staging -> device)
struct TdSelectParams {
  int use_overlap;
  int is_seam_image;         // rows / cols match a merged strip: no overlap-band discard
  double left, bottom, right, top;                          // raster bounds (ndvi_bounds)
  double band_left, band_right, band_top, band_bottom;      // overlap-band borders
  float height_threshold, ndvi_mean_threshold, ndvi_var_threshold;
};

struct TdAffine6 {
  double a, b, c, d, e, f;
};

struct TdImageParams {
  long long n_inst;          // live ROI-head instances of this image
  TdAffine6 ndvi_tf;         // transform of the (decimated) NDVI raster
  TdAffine6 height_tf;       // transform of the (decimated) height raster
  TdSelectParams sel;
};

// P2: n_dev != null -> n_inst is the capacity, rows >= *n_dev become empty windows
int td_paste_plan_ex(const float* boxes_net, const int* inst_tile, const int* tile_dims, int n_inst, int n_tiles,
                     float* boxes_px, int* win, long long* nwords, long long* npx, const long long* n_dev,
                     cudaStream_t st);

// td_take_rings with the coordinate rounding of P9 (round_coordinates) fused into the copy
int td_take_rings_ex(const double* verts, const long long* ring_off, const long long* sel, int n_out,
                     const int* scratch, const long long* dst_off, double* out_verts, const long long* n_dev,
                     int round3, cudaStream_t st);

// P7 with a ring indirection (crown k = ring ring_idx[k] of the table) and a device-side transform
int td_crown_stats_ex(const double* verts, const long long* ring_off, const long long* ring_idx, int n,
                      const float* ndvi, const float* height, int rows, int cols, const TdAffine6* tf_host,
                      const TdAffine6* tf_dev, int mode, float* max_h, float* hxy, float* ndvi_stats,
                      const long long* n_dev, cudaStream_t st);
int td_centroids_ex(const double* verts, const long long* ring_off, const long long* ring_idx, int n, float* centroid,
                    int* vmax_scratch, const long long* n_dev, cudaStream_t st);

// P8 on float64 bounds (cast to float32 like the reference's cp.array(..., dtype=float32))
int td_containment_ex(const double* bounds64, const float* bounds32, int n, double threshold, float* ratio_max,
                      unsigned char* is_contained, int* num_contained, const long long* n_dev, cudaStream_t st);

// P9 selection with the parameters in device memory
int td_select_crowns_ex(const double* bounds, const float* max_h, const float* ndvi_stats, const double* area,
                        const int* num_contained, const unsigned char* is_contained, int n,
                        const TdSelectParams* p_host, const TdSelectParams* p_dev, int* pre, int* out_idx,
                        const long long* n_dev, cudaStream_t st);

// compaction of the indices whose flag equals `want` (0 or 1)
int td_compact_flags_ex(const unsigned char* flags, int want, int n, const long long* n_dev, long long* sel,
                        long long* count, cudaStream_t st);
