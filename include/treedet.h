/* libtreedet -- C-ABI of the B200-native crown pipeline (sm_100a).
 *
 * Drop-in boundary for the post-model path of Jonetz/TreeDetection.  The reference has
 * no FFI of its own: the path is Python calling CuPy / detectron2 / rasterio / shapely
 * (SURVEY.md section 8b).  Each entry point below names the reference function(s)
 * (file:line under /root/reference) whose work it replaces; INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add at that call site.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the comment says "host";
 *  - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream);
 *    calls are asynchronous on that stream unless stated otherwise;
 *  - outputs are caller allocated; temporary scratch comes from the stream-ordered
 *    CUDA memory pool (cudaMallocAsync) and is released before returning; no global
 *    state, so calls on different streams / devices are independent;
 *  - return value: 0 = ok, <0 = error (TD_ERR_*), text via td_last_error() (thread local);
 *  - ragged polygon rings: `verts` (V,2) float64 interleaved x,y + `ring_off` (R+1) int64;
 *    rings are closed (first vertex repeated at the end);
 *  - `n_dev` (where present, may be null): DEVICE pointer to the live item count; the by-value
 *    count is then the capacity the grid is sized for and items >= *n_dev are left untouched.
 *    This is what lets a whole image run without a host synchronisation (see "Device-side
 *    bookkeeping" below);
 *  - there is no CPU fallback anywhere in this library.
 */
#ifndef TREEDET_H_
#define TREEDET_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TD_OK 0
#define TD_ERR_CUDA (-1)
#define TD_ERR_ARG (-2)
#define TD_ERR_OVERFLOW (-3)
#define TD_ERR_UNSUPPORTED (-4)

int td_version(void);               /* 100 = 0.1.0 */
const char* td_last_error(void);    /* host string, thread local */
int td_device_sms(void);            /* SM count of the current device (148 on B200) */

/* ---- P1: tile cut + normalise ------------------------------------------------------
 * Replaces Predictor._process_tile (TreeDetection/prediction.py:159-176): rasterio.mask
 * crop of the tile window, band reorder (2,1,0), optional 255*x/65535 for 16-bit data,
 * detectron2 ResizeShortestEdge(800, 1333) (PIL bilinear for uint8), float32 CHW.
 * Plan / execute: the tile tables (windows, PIL coefficient tables) are built and uploaded
 * once per tiling, every image with that tiling re-uses the plan.
 *   tile_win   HOST (T,4) int32 [col_off, row_off, w, h]   (tiling.tile_grid)
 *   tile_net   HOST (T,2) int32 [net_h, net_w]             (tiling.resize_shortest_edge)
 *   out_off    HOST (T+1) int64 float offsets of each tile's (3, net_h, net_w) block in `out`
 *   elem_size  1 = uint8 raster, 2 = uint16 raster; H, W = raster size                    */
int td_tile_plan_create(const int* tile_win, const int* tile_net, const long long* out_off, int n_tiles,
                        int elem_size, int H, int W, void** plan_out);
int td_tile_plan_destroy(void* plan);
/*   image      (bands, H, W) planar raster on the device (16-byte aligned `out`)
 *   rescale16  (T) uint8 out, may be null: 0 = uint8 tile, 1 = 16-bit branch taken
 *              (max(band 1) > 255), 2 = uint16 tile the reference fails on (left untouched) */
int td_tile_cut_normalize(const void* plan, const void* image, int bands, float* out, unsigned char* rescale16,
                          void* stream);

/* ---- P2: mask paste + threshold + bit-pack -------------------------------------------
 * Replaces detectron2 detector_postprocess + paste_masks_in_image + _do_paste_mask
 * (entered at TreeDetection/prediction.py:181-183) and the identity resize + uint8 cast
 * of prediction.py:222-229.
 *   boxes_net (N,4) f32 xyxy in network-input pixels; inst_tile (N) i32;
 *   tile_dims (T,4) i32 [tile_h, tile_w, net_h, net_w]
 *   -> boxes_px (N,4) f32 (scaled, clipped), win (N,4) i32 [x0,y0,w,h] (0,0,0,0 when the
 *      box is empty and the instance is dropped), nwords (N) i64 = ceil(w/32)*h      */
int td_paste_plan(const float* boxes_net, const int* inst_tile, const int* tile_dims, int n_inst, int n_tiles,
                  float* boxes_px, int* win, long long* nwords, long long* npx, void* stream);
/*   npx (2,N) i64, may be null: row 0 = w*h of every window (label plane of td_trace_emit / _walk),
 *   row 1 = the instance's point slot for td_trace_walk, 4*(w+h)+64 (0 for dropped instances)     */
/*   word_off (N+1) i64 = exclusive scan of nwords; probs (N,28,28) f32 probabilities;
 *   bits: packed 1-bit rasters, row-major, 32 pixels per uint32 (LSB = leftmost)       */
int td_paste_threshold_pack(const float* boxes_px, const int* win, const long long* word_off, const float* probs,
                            int n_inst, float threshold, uint32_t* bits, void* stream);
/*   the pasted float32 probabilities themselves (tolerance tests): val_off = scan of w*h */
int td_paste_values(const float* boxes_px, const int* win, const long long* val_off, const float* probs,
                    int n_inst, float* vals, void* stream);

/* ---- P3: border following -> CRS rings ---------------------------------------------------
 * Replaces Predictor._process_and_save_single (TreeDetection/prediction.py:197-265:
 * cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE), contour.size >= 8, closing point) and
 * xy_gpu (TreeDetection/utilities.py:182-207).  Two passes (sizes are data dependent).
 *   planes  scratch, 2*total_words uint32 (zeroed by the call)
 *   counts  (N,4) i32 out: [borders, points, kept rings, ring vertices]; borders < 0 when
 *           one window holds more than 65534 borders                                      */
int td_trace_count(const uint32_t* bits, const int* win, const long long* word_off, int n_inst,
                   long long total_words, uint32_t* planes, int* counts, long long* sizes_kn, void* stream);
/*   sizes_kn (4,N) i64, may be null: the same counts as rows (the layout td_scan_clamp consumes)  */
/*   labels (sum w*h) u16 scratch; px_off / cont_off / pts_off / ring_base / vert_base:
 *   (N+1) i64 exclusive scans of w*h and of the four count columns; ct_int: 6*total_contours
 *   i32 scratch; ct_hole: total_contours u8 scratch; pts: 2*total_points i16 scratch;
 *   tile_tf (T,6) f64 window transforms.  Outputs: ring_off[0..R) (caller sets
 *   ring_off[R] = V), ring_inst (R) i32 producing instance, verts (V,2) f64.
 *   Ring order = tile-major instance order, then cv2's contour order.                     */
int td_trace_emit(const uint32_t* bits, const int* win, const long long* word_off, int n_inst,
                  long long total_words, uint32_t* planes, unsigned short* labels, const long long* px_off,
                  const long long* cont_off, const long long* pts_off, const long long* ring_base,
                  const long long* vert_base, int* ct_int, unsigned char* ct_hole, short* pts,
                  long long total_contours, const int* inst_tile, const double* tile_tf, long long* ring_off,
                  int* ring_inst, double* verts, void* stream);

/*   Single-pass form for the sync-free chain: the walk writes into per-instance slots (cap_contours
 *   table rows at i*cap_contours, points at pts_off[i]..pts_off[i+1]) and counts at the same time;
 *   an instance that outgrows its slot raises bit 2 of *flag.  ct_int6: 6*N*cap_contours i32,
 *   ct_hole: N*cap_contours u8, pts: 2*pts_off[N] i16, counts (N,4) i32, sizes_kn (2,N) i64 = [kept
 *   rings, ring vertices].  td_trace_rings turns the slots into the outputs of td_trace_emit given
 *   ring_base / vert_base = exclusive scans of sizes_kn.                                          */
int td_trace_walk(const uint32_t* bits, const int* win, const long long* word_off, int n_inst,
                  long long total_words, uint32_t* planes, unsigned short* labels, const long long* px_off,
                  const long long* pts_off, int cap_contours, int* ct_int6, unsigned char* ct_hole, short* pts,
                  int* counts, long long* sizes_kn, long long* flag, void* stream);
int td_trace_rings(const int* win, int n_inst, const int* counts, const long long* pts_off, int cap_contours,
                   int* ct_int6, const unsigned char* ct_hole, const short* pts, const long long* ring_base,
                   const long long* vert_base, const int* inst_tile, const double* tile_tf, long long* ring_off,
                   int* ring_inst, double* verts, void* stream);

/* ---- P4 (+ the area of P9's head): simplify, tile box filter ---------------------------------
 * Replaces process_prediction_file_sync (TreeDetection/helpers.py:419-476: shapely
 * simplify(tol, preserve_topology=True) + sjoin "within" the shrunk tile box of
 * box_make, helpers.py:280-303) and shape(geom).simplify(2).area
 * (TreeDetection/postprocessing.py:747-754).
 *   scratch 5*V i32 (kept-vertex index lists live at scratch[5*ring_off[r] ..]);
 *   alive   (V/32 + R + 2) u32 scratch;
 *   boxes (B,4) f64 + ring_box (R) i32, or both null (no filter);
 *   out_count (R) i32 kept vertices; out_bounds (R,4) f64 / out_area (R) f64 / out_keep (R)
 *   u8 may be null.  tolerance <= 0: no simplification (bounds / area of the ring itself).
 *   bounds_of_input != 0: out_bounds holds the bounds of the INPUT ring (polygon.bounds of the
 *   un-simplified crown, postprocessing.py:497-503) while out_area is that of simplify(tol).  */
int td_simplify_rings(const double* verts, const long long* ring_off, int n_rings, double tolerance, int* scratch,
                      uint32_t* alive, const double* boxes, const int* ring_box, int* out_count,
                      double* out_bounds, double* out_area, unsigned char* out_keep, int bounds_of_input,
                      const long long* n_dev, void* stream);
/*   output ring q = input ring sel[q]; dst_off (n_out+1) i64; scratch = the index lists
 *   above (copy kept vertices only) or null (copy whole rings)                           */
int td_take_rings(const double* verts, const long long* ring_off, const long long* sel, int n_out,
                  const int* scratch, const long long* dst_off, double* out_verts, const long long* n_dev,
                  void* stream);

/* ---- P5: decimated raster reads + NDVI ----------------------------------------------------
 * Replaces the rasterio reads with out_shape + Resampling.bilinear in process_geojson
 * (TreeDetection/postprocessing.py:780-800) and ndvi_array_from_rgbi / ndvi_index
 * (TreeDetection/helpers.py:862-896).  rgbi (bands>=4, H, W) planar uint8.            */
int td_ndvi_decimate(const unsigned char* rgbi, int bands, int in_h, int in_w, int out_h, int out_w,
                     float* ndvi_out, void* stream);
int td_decimate_f32(const float* src, int in_h, int in_w, int out_h, int out_w, float* out, void* stream);

/* ---- P6: ordered bbox NMS -------------------------------------------------------------------
 * Replaces filter_polygons_by_iou_and_area (TreeDetection/postprocessing.py:349-406) and
 * calculate_iou (TreeDetection/utilities.py:112-144).  bounds (N,4) f64, conf / area (N)
 * f64 (cast to float32 / float16 / float16 as the reference does); removed (N) u8.
 * Synchronises the stream once (one 8-byte read back sizes the adjacency).            */
int td_bbox_nms_ordered(const double* bounds, const double* conf, const double* area, int n,
                        double iou_threshold, double area_threshold, unsigned char* removed, void* stream);
/*   capacity form (no synchronisation): n = capacity, *n_dev = live count, neighbour lists limited
 *   to nbr_cap entries; bit 1 of *flag is raised when nbr_cap is too small (removed undefined) and
 *   nothing is resolved when *flag is already non-zero on entry                                  */
int td_bbox_nms_ordered_dyn(const double* bounds, const double* conf, const double* area, int n,
                            const long long* n_dev, double iou_threshold, double area_threshold, long long nbr_cap,
                            long long* flag, unsigned char* removed, void* stream);

/*   Opt-in mask-IoU cleaner on the packed rasters of P2 (``iou_mode: mask``): the rule of clean_crowns
 *   (TreeDetection/helpers.py:602-701, never called by the reference) with the PIXEL IoU
 *   popcount(a & b) / popcount(a | b) of the 1-bit crown rasters on the image's pixel grid.
 *   win (N,4) i32 [x0,y0,w,h] in tile pixels, tile_org (T,2) i32 [col_off,row_off] of each tile window,
 *   scores (N) f32 -> keep (N) u8, match (N) i32 (best-confidence crown with IoU > iou_thr, -1: none),
 *   best_iou (N) f32 nullable                                                                      */
int td_mask_iou_clean(const uint32_t* bits, const long long* word_off, const int* win, const int* tile_org,
                      const int* inst_tile, const float* scores, int n, float iou_thr, float confidence,
                      unsigned char* keep, int* match, float* best_iou, void* stream);

/* ---- P7: per-crown raster statistics, centroids ----------------------------------------------
 * Replaces get_metadata_within_polygon (TreeDetection/postprocessing.py:221-347; mode 0),
 * get_height_within_polygon (:25-115; mode 1), get_ndvi_within_polygon (:117-219; mode 2),
 * is_point_in_polygon_batch (TreeDetection/utilities.py:78-98) and get_centroids
 * (utilities.py:163-180).  transform6: HOST pointer to (a,b,c,d,e,f).
 *   max_h (N) f32, hxy (N,2) f32, ndvi_stats (N,4) f32 [min,max,mean,var]; -1 when empty. */
int td_crown_stats(const double* verts, const long long* ring_off, int n, const float* ndvi, const float* height,
                   int rows, int cols, const double* transform6, int mode, float* max_h, float* hxy,
                   float* ndvi_stats, const long long* n_dev, void* stream);
int td_centroids(const double* verts, const long long* ring_off, int n, float* centroid, const long long* n_dev,
                 void* stream);
/*   optional nDSM summary (north_star "min / max / mean / percentile"; the reference keeps the maximum only,
 *   postprocessing.py:25-115): over the pixel set of get_height_within_polygon, out (N,4) f32 = [min, mean,
 *   percentile q (0..100, numpy's "linear" rule, exact order statistics), pixel count]; empty set -> -1   */
int td_crown_height_summary(const double* verts, const long long* ring_off, int n, const float* height, int rows,
                            int cols, const double* transform6, double q, float* out, const long long* n_dev,
                            void* stream);

/* ---- P8: bbox containment ---------------------------------------------------------------------
 * Replaces process_containment_features (TreeDetection/postprocessing.py:408-476).
 * bounds32 (N,4) f32 -> ratio_max (N) f32, is_contained (N) u8, num_contained (N) i32.  */
int td_containment(const float* bounds32, int n, double threshold, float* ratio_max, unsigned char* is_contained,
                   int* num_contained, const long long* n_dev, void* stream);

/* ---- P9: selection + coordinate rounding ----------------------------------------------------
 * Replaces the pre-selection and containment case analysis of process_features
 * (TreeDetection/postprocessing.py:571-667), element_is_near_border
 * (TreeDetection/helpers.py:501-522) and round_coordinates (utilities.py:146-161).
 *   params: HOST pointer to 14 doubles [use_overlap, is_seam_image, left, bottom, right, top,
 *   band_left, band_right, band_top, band_bottom, height_thr, ndvi_mean_thr, ndvi_var_thr, 0]
 *   pre (N) i32: pre-selected flag; out_idx (N) i32: crown emitted by crown i, or -1.    */
int td_select_crowns(const double* bounds, const float* max_h, const float* ndvi_stats, const double* area,
                     const int* num_contained, const unsigned char* is_contained, int n, const double* params,
                     int* pre, int* out_idx, const long long* n_dev, void* stream);
int td_round_coords(const double* in, long long n, double* out, void* stream);
/*   head of process_geojson (postprocessing.py:739-768): flags (N) u8 = conf >= conf_thr and
 *   area_min <= area <= area_max (area of simplify(2)); poly_id (N) i64 = enumeration index after the
 *   confidence filter                                                                              */
int td_select_head(const double* conf, const double* area, int n, const long long* n_dev, double conf_thr,
                   double area_min, double area_max, unsigned char* flags, long long* poly_id, void* stream);

/* ---- P10: forest-outline predicates (two-model fusion, tile flags) -----------------------------
 * Replaces the GEOS predicates of fuse_predictions (TreeDetection/helpers.py:795-811) and of
 * tile_single_file (TreeDetection/preprocessing.py:67-96).  a_*: query rings (crowns or tile
 * boxes); f_verts / f_off: forest RINGS; f_poly_off (n_poly + 1) i64 groups them into polygons
 * (first ring = shell, the others = holes; null: every ring is a polygon without holes);
 * f_bounds (n_poly,4) f64 = bounds of each polygon's shell; a_filter (n_a,4) f64 or null: pick
 * candidate polygons by strict bbox overlap with this box (the un-buffered tile box) instead of the
 * ring's own bounds.
 * out_intersects / out_within (n_a) u8: 1 = ring intersects / lies within the union of the
 * forest polygons, 0 = not; out_within 2 = more than 62 crossings on one edge of the query.  */
int td_forest_predicates(const double* a_verts, const long long* a_off, int n_a, const double* f_verts,
                         const long long* f_off, const long long* f_poly_off, const double* f_bounds, int n_poly,
                         const double* a_filter, unsigned char* out_intersects, unsigned char* out_within,
                         void* stream);

/*   out (n) u8 = ring r is a valid polygon shell (simple closed ring): the `is_valid` test that selects the
 *   geometries fuse_predictions repairs with buffer(0) / make_valid (TreeDetection/helpers.py:816-821)   */
int td_ring_is_simple(const double* verts, const long long* ring_off, int n_rings, unsigned char* out, void* stream);

/* ---- Device-side bookkeeping of the sync-free chain ----------------------------------------------
 * The reference sizes every intermediate through the host (len(), .get(), Python lists, e.g.
 * TreeDetection/postprocessing.py:389-405, 739-768).  These helpers keep counts on the device.
 *   td_scan_clamp: sizes (k,n) i64 -> offs (k,n+1) i64 exclusive scans, truncated at the first item
 *     whose end exceeds caps[r] (HOST array of k capacities) in any row or whose size is negative:
 *     that item and all later ones become empty, bit 0 of *flag is raised, win_zero (nullable, (n,4)
 *     i32) gets w = h = 0 for them; totals (k) i64 = offs[r][n].
 *   td_compact_flags: sel[0..count) = ascending i with flags[i] != 0 (i < n, i < *n_dev), tail 0.
 *   td_compact_nonneg: out[0..count) = the non-negative values[i] in order, tail 0.
 *   td_ring_tail: ring_off[i] = *n_verts for *n_rings <= i <= cap_rings, ring_inst tail = 0.        */
int td_scan_clamp(const long long* sizes, int k, int n, const long long* caps, long long* offs, long long* totals,
                  long long* flag, int* win_zero, void* stream);
int td_compact_flags(const unsigned char* flags, int n, const long long* n_dev, long long* sel, long long* count,
                     void* stream);
int td_compact_nonneg(const int* values, int n, const long long* n_dev, long long* out, long long* count,
                      void* stream);
int td_ring_tail(long long* ring_off, int* ring_inst, int cap_rings, const long long* n_rings,
                 const long long* n_verts, void* stream);
/*   td_ring_offsets: dst_off (n+1) i64 = exclusive offsets of the rings sel[0..n) -- lengths from
 *     `count` (kept vertices, td_simplify_rings) when given, else from ring_off.
 *   td_gather_rows: k <= 8 row gathers in one launch, out[a][i] = in[a][sel[i]] for i < n (< *n_dev);
 *     in / out / row_bytes are HOST arrays of k device pointers / row sizes in bytes.               */
int td_ring_offsets(const long long* ring_off, const int* count, const long long* sel, int n, long long* dst_off,
                    void* stream);
int td_gather_rows(const void* const* in, void* const* out, const int* row_bytes, int k, const long long* sel, int n,
                   const long long* n_dev, void* stream);

/* ---- The whole chain of one image as one host call per stage (CUDA-graph replay) -------------------
 * Replaces, for a stream of images with the same tiling, the per-image host loops of
 *   Predictor._process_and_save_single + process_and_stitch_predictions
 *       (TreeDetection/prediction.py:197-265, TreeDetection/helpers.py:419-600)   -> td_chain_predict
 *   process_geojson + process_features (TreeDetection/postprocessing.py:722-809, 478-720) -> td_chain_post
 * Every variable-length intermediate lives, by capacity, in ONE caller-allocated device workspace; live
 * counts stay on the device (16 int64 counters per output slot: [overflow flag, words, label pixels,
 * point slots, -, -, traced rings, traced vertices, table rings, table vertices, after head, after NMS,
 * final crowns, final vertices, -, -]); the static launch sequence is captured into a CUDA graph per
 * (slot, input pointers) and replayed with one cudaGraphLaunch.  An overflow of a capacity raises a bit
 * of counter 0 (1 = a buffer, 2 = NMS neighbour slots, 4 = contour slot of the single-pass walk) and
 * the image has to be redone through the exact-size entry points above.
 *   cfg   HOST 16 doubles: [mask_threshold, simplify_tolerance, area simplify tolerance (2.0),
 *         confidence_threshold, area_min, area_max, iou_threshold, area_threshold,
 *         containment_threshold, 0 ...]
 *   caps  HOST 8 int64: [instances, packed words, label pixels, point slots, rings, vertices, NMS
 *         neighbour slots per crown, contour rows per instance]
 *   workspace: device memory, 256-byte aligned, td_chain_workspace_bytes(caps, n_slots) bytes
 *   n_slots (1..8): output slots (results of slot s stay valid until slot s is used again)        */
long long td_chain_workspace_bytes(const long long* caps, int n_slots);
int td_chain_create(const double* cfg, const long long* caps, int n_slots, void* workspace,
                    long long workspace_bytes, void** chain_out);
int td_chain_destroy(void* chain);
/*   byte offsets (HOST, 16 int64 out) of slot `slot`'s outputs inside the workspace: counters (16 i64),
 *   table verts (V,2) f64, table ring_off (R+1) i64, table conf (R) f64, verts (V,2) f64 rounded,
 *   ring_off (R+1) i64, poly_id (R) i64, conf (R) f64, area (R) f64, tree_height (R) f32, centroid (R,2)
 *   f32, is_contained (R) u8, num_contained (R) i32, height arg-max xy (R,2) f32, ndvi stats (R,4) f32  */
int td_chain_layout(const void* chain, int slot, long long* offsets16);
/*   P2 + P3 + P4: boxes_net (N,4) f32, scores (N) f32, probs (N,28,28) f32, inst_tile (N) i32,
 *   tile_dims (T,4) i32, tile_tf (T,6) f64, tile_boxes (T,4) f64 (helpers.py:280-303), all device      */
int td_chain_predict(void* chain, int slot, const float* boxes_net, const float* scores, const float* probs,
                     const int* inst_tile, int n_inst, const int* tile_dims, const double* tile_tf,
                     const double* tile_boxes, int n_tiles, int use_graph, void* stream);
/*   P9 head + P6 + P7 + P8 + P9 on the table of `slot`: ndvi / height device rasters (f32) with HOST
 *   6-double transforms; combined != 0: get_metadata_within_polygon (shared grid), else the split pair;
 *   select_params: the 14 HOST doubles of td_select_crowns                                         */
int td_chain_post(void* chain, int slot, const float* ndvi, int ndvi_rows, int ndvi_cols, const double* ndvi_tf,
                  const float* height, int height_rows, int height_cols, const double* height_tf, int combined,
                  const double* select_params, int use_graph, void* stream);

/* ---- P0a: seam strips ---------------------------------------------------------------------------
 * Replaces crop_single_image / merge_images / crop_image (TreeDetection/merging.py:34-110,
 * TreeDetection/helpers.py:1023-1085): mosaic of an image with its right (axis 0) or lower
 * (axis 1) neighbour, centre-cropped to strip_w x strip_h pixels, without building the mosaic.
 *   a, b: (bands, H, W) planar rasters of elem_size bytes; out (bands, strip_h, strip_w).  */
int td_seam_crop(const void* a, const void* b, int elem_size, int bands, int ha, int wa, int hb, int wb, int axis,
                 int strip_w, int strip_h, void* out, void* stream);

/* ---- N1: GeoTIFF LZW codec (HOST function, host pointers) ------------------------------------------
 * Replaces the GDAL LZW decoder behind rasterio.open(...).read (TreeDetection/prediction.py:61,
 * postprocessing.py:781-800, merging.py:56-75).  One call decodes one strip / tile (TIFF 6.0
 * section 13); returns the bytes written to dst (<= cap) or a negative TD_ERR_* code.            */
long long td_tiff_lzw_decode(const unsigned char* src, long long n_src, unsigned char* dst, long long cap);

/* The encoder of the same format (HOST function): ClearCode, MSB-first codes with early change, ClearCode when
 * the table is full, EOI -- what libtiff writes.  Returns the bytes written (cap >= n_src * 3 / 2 + 16 suffices) or
 * a negative TD_ERR_* code.  Used by this package's GeoTIFF writer (the reference writes its merged seam strips
 * through rasterio: TreeDetection/merging.py:100-107).                                                       */
long long td_tiff_lzw_encode(const unsigned char* src, long long n_src, unsigned char* dst, long long cap);

/* The same decoder on the DEVICE, for all strips / tiles of a raster at once (device pointers): chunk k is the
 * LZW stream src[src_pos[k] : src_pos[k] + src_len[k]] -- src is typically the whole file, copied to the device
 * still compressed -- and is decoded to dst + k * dst_stride (dst_len[k] expected bytes, dst_stride <= 1 MiB; a
 * shorter stream is zero filled).  out_len (n_chunks, nullable) receives the decoded byte counts; *status (one
 * int) becomes a TD_ERR_* code if any stream is corrupt or overflows its chunk.  One warp per stream.       */
int td_tiff_lzw_decode_batch(const unsigned char* src, const long long* src_pos, const int* src_len, int n_chunks,
                             unsigned char* dst, long long dst_stride, const int* dst_len, int* out_len, int* status,
                             void* stream);

/* Decoded chunks -> the planar (bands, height, width) raster the path consumes: undoes TIFF predictor 2
 * (horizontal differencing, 8-bit samples) and de-interleaves chunky pixels.  Replaces what GDAL does behind
 * rasterio's read() after the codec (TreeDetection/prediction.py:61, postprocessing.py:781-800).
 *   decoded: chunk k at decoded + k * stride, in TIFF order (planar == 2: plane-major);
 *   sample_size 1 (predictor 1 or 2) or 4 (predictor 1); planar = TIFF PlanarConfiguration (1 chunky, 2 planar). */
int td_tiff_place_chunks(const unsigned char* decoded, long long stride, int n_chunks, void* out, int bands, int height,
                         int width, int sample_size, int planar, int chunk_rows, int chunk_cols, int predictor,
                         void* stream);

/* ---- N2: GeoPackage feature writer (HOST function, host pointers) ----------------------------------
 * Replaces the row loop of GeoDataFrame.to_file(driver="GPKG") behind the stitched layer
 * (TreeDetection/helpers.py:592-599) and the processed layer (postprocessing.py:903-936): appends n_rings
 * polygon features (GeoPackageBinary: header, envelope, WKB polygon with one ring) to table `layer` of an
 * existing GeoPackage inside one transaction.
 *   verts (V,2) f64, ring_off (n_rings + 1) i64;
 *   col_types[c]: 0 = float64 array (NaN -> NULL), 1 = int64 array, 2 = text (col_data[c] = UTF-8 bytes,
 *   col_text_off[c] = n_rings + 1 byte offsets).  SQLite is loaded with dlopen("libsqlite3.so.0"):
 *   TD_ERR_UNSUPPORTED when it is not there.                                                          */
int td_gpkg_append(const char* path, const char* layer, int epsg, const double* verts, const long long* ring_off,
                   long long n_rings, int n_cols, const char* const* col_names, const int* col_types,
                   const void* const* col_data, const long long* const* col_text_off);

#ifdef __cplusplus
}
#endif
#endif /* TREEDET_H_ */
