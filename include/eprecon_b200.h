/* eprecon_b200 — C ABI of the B200 (sm_100a) feature-volume hot path of zhen6618/EPRecon.
 *
 * The reference has no FFI layer: its boundary is Python nn.Module.forward().  The
 * drop-in modules in eprecon_b200/*.py keep those signatures and call the entry points below through ctypes
 * with tensor.data_ptr() and the current cudaStream_t.  Every function
 *   - takes only raw device pointers, sizes and a stream (no torch types),
 *   - returns 0 or a negative status (EP_ERR_*), never throws, never allocates (callers pass workspaces sized
 *     by the matching *_workspace_bytes query), never synchronises,
 *   - is deterministic (no floating-point atomics).
 * Each block names the reference code it replaces (file:line under the reference tree).
 */
#ifndef EPRECON_B200_H_
#define EPRECON_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_API_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define EP_OK 0
#define EP_ERR_ARG (-1)
#define EP_ERR_WORKSPACE (-2)
#define EP_ERR_CUDA (-3)
#define EP_ERR_UNSUPPORTED (-4)

int ep_version(void);
/* host wait policy of stream synchronisation on the current device: 0 default, 1 spin, 2 yield, 4 blocking (csrc/capi.cu) */
int ep_set_sync_mode(int mode);

/* ---- multi-view back-projection ---------------------------------------------------------------------------
 * replaces models/occupancy_initialization.py:189-261 (Back_Project.forward), :79-128 (init-stage projection +
 * masked variance) and ops/back_project.py:5-80 (legacy back_project).
 * coords int32 [n,4]=(b,x,y,z); origin f32 [bs,3]; krcam f32 [V,bs,4,4]; feats channels-last f32 [V,bs,H,W,C]. */
size_t ep_backproject_workspace_bytes(int64_t n);
int ep_backproject_count(const int32_t* coords, int64_t n, const float* origin, float voxel_size,
                         const float* krcam, int n_views, int bs, int feat_h, int feat_w, int min_views,
                         float* count, uint32_t* vismask, int32_t* valid_per_batch, int32_t* n_valid_total,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream);
int ep_backproject_compact(const int32_t* coords, const uint32_t* vismask, int64_t n, int min_views,
                           int32_t* out_coords, uint32_t* out_vis, int32_t* out_src, const void* workspace,
                           cudaStream_t stream);
/* mode 0: masked mean over views; mode 1: masked two-pass population variance.  zbar (optional): mean depth. */
int ep_backproject_gather(const int32_t* out_coords, const uint32_t* out_vis, int64_t m, const float* feats_nhwc,
                          int channels, int n_views, int bs, int feat_h, int feat_w, const float* origin,
                          float voxel_size, const float* krcam, int mode, float* out, int ld_out, float* zbar,
                          cudaStream_t stream);
/* single-pass variant (count + stable compaction by decoupled look-back + gather in one kernel); outputs have capacity n,
 * totals[bs+1] = survivors per batch entry and the grand total.  channels in {24,32,40,80}. */
size_t ep_backproject_fused_workspace_bytes(int64_t n);
int ep_backproject_fused(const int32_t* coords, int64_t n, const float* origin, float voxel_size, const float* krcam,
                         int n_views, int bs, int feat_h, int feat_w, const float* feats_nhwc, int channels,
                         int min_views, int mode, float* count, int32_t* out_coords, uint32_t* out_vis,
                         int32_t* out_src, float* out, int ld_out, float* zbar, int32_t* totals, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream);
int ep_backproject_grid(const int32_t* out_coords, const uint32_t* out_vis, int64_t m, int n_views, int bs,
                        int feat_h, int feat_w, const float* origin, float voxel_size, const float* krcam,
                        float* im_grid, uint8_t* mask, cudaStream_t stream);
int ep_nchw_to_nhwc(const float* in, float* out, int n_img, int channels, int hw, cudaStream_t stream);
/* input side (models/neuralrecon.py:53-54, neucon_network.py:364): V separately allocated per-view NCHW maps [bs,C,hw] (host
 * array of device pointers) -> one channels-last buffer [V,bs,hw,C]; replaces torch.stack + ep_nchw_to_nhwc */
int ep_pack_views_nhwc(const float* const* views, int n_views, int bs, int channels, int hw, float* out, cudaStream_t stream);

/* ---- scan / compaction / sort plumbing ----------------------------------------------------------------------
 * replaces torch.nonzero / boolean indexing / torch.unique (models/neucon_network.py:304,312,492-501;
 * ops/torchsparse_utils.py:20; utils.py:172,179). */
size_t ep_compact_workspace_bytes(int64_t n);
int ep_compact_flags(const uint8_t* flags, int64_t n, int32_t* out_index, int32_t* out_pos, int32_t* total_dev,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t ep_sort_segments_workspace_bytes(int64_t n);
int ep_sort_segments(const uint64_t* keys, int64_t n, int key_bits, uint64_t sentinel, uint64_t* keys_sorted,
                     int32_t* perm, int32_t* seg_start, int32_t* seg_end, int32_t* seg_of_item,
                     int32_t* n_segments_dev, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- coordinate hashing, tables, kernel maps ----------------------------------------------------------------
 * replaces torchsparse v2.0.0 sphash / sphashquery / spdownsample / kernel-map search / calc_ti_weights
 * (ops/torchsparse_utils.py:15-105; inside spnn.Conv3d, models/modules.py:19-181) and spconv indice pairs
 * (models/modules.py:252,444). */
int ep_hash_build(const uint64_t* keys, int64_t m, uint64_t* table_keys, int32_t* table_vals, int64_t capacity,
                  cudaStream_t stream);
int ep_coord_keys(const int32_t* coords, int64_t m, int batch_first, uint64_t* keys, cudaStream_t stream);
int ep_point_keys(const float* pts, int64_t n, float vres, int spatial, float* pts_scaled, uint64_t* keys,
                  cudaStream_t stream);
int ep_segment_coords(const float* pts_scaled, const int32_t* perm, const int32_t* seg_start, int64_t m,
                      int32_t* vox_coords, cudaStream_t stream);
int ep_kmap_build(const int32_t* out_coords, int64_t m_out, int batch_first, const int32_t* offsets, int K,
                  const uint64_t* table_keys, const int32_t* table_vals, int64_t capacity, int sx, int sy, int sz,
                  int32_t* nbr, cudaStream_t stream);
int ep_kmap_inverse(const int32_t* nbr, int64_t m_out, int K, int32_t* inv, cudaStream_t stream);
int ep_down_keys(const int32_t* coords, int64_t m, int step, uint64_t* keys, cudaStream_t stream);
int ep_down_unpack(const uint64_t* keys_sorted, const int32_t* seg_start, int64_t m, int32_t* coords,
                   cudaStream_t stream);
/* compact Z-order keys: fewer radix-sort passes, same row order (csrc/common.cuh); *violation != 0 -> redo with the wide form */
int ep_point_keys_compact(const float* pts, int64_t n, float vres, int coord_bits, int batch_bits, float* pts_scaled,
                          uint64_t* keys, int32_t* violation, cudaStream_t stream);
int ep_down_keys_compact(const int32_t* coords, int64_t m, int step, int coord_bits, uint64_t* keys, cudaStream_t stream);
int ep_down_unpack_compact(const uint64_t* keys_sorted, const int32_t* seg_start, int64_t m, int coord_bits, int32_t* coords,
                           cudaStream_t stream);
int ep_devox_prepare(const float* pts_scaled, int64_t n, int stride, const uint64_t* table_keys,
                     const int32_t* table_vals, int64_t capacity, int32_t* idx, float* weights, cudaStream_t stream);
int ep_point_query(const float* pts_scaled, int64_t n, int stride, const uint64_t* table_keys,
                   const int32_t* table_vals, int64_t capacity, int32_t* idx, uint64_t* idx_as_key, int m_sentinel,
                   cudaStream_t stream);

/* ---- point <-> voxel feature transfer -----------------------------------------------------------------------
 * replaces torchsparse spvoxelize / spdevoxelize (ops/torchsparse_utils.py:27,58,86,98). */
int ep_segment_mean(const float* feat, int ld_in, int c, const int32_t* perm, const int32_t* seg_start,
                    const int32_t* seg_end, int64_t m, float* out, int ld_out, cudaStream_t stream);
int ep_devoxelize(const float* feat, int ld_in, int c, const int32_t* idx, const float* weights, int64_t n,
                  const float* add, int ld_add, float* out, int ld_out, cudaStream_t stream);

/* ---- sparse convolution / linear / batch-norm statistics ----------------------------------------------------
 * replaces spnn.Conv3d forward, spconv SubMConv3d, nn.Linear and the statistics pass of train-mode BatchNorm
 * (models/modules.py:15-73,89-136,178-197,249-311,401-452; main.py:357). */
int ep_spconv_num_row_tiles(int64_t m_out);
int ep_spconv_fwd(const float* in, int ld_in, int cin, const int32_t* nbr, int K, const float* W, int ldw, int cout,
                  const float* bias, float* out, int ld_out, int64_t m_out, float* bn_partial, cudaStream_t stream);
/* tensor-core variant: tcgen05.mma kind::tf32, fp32 accumulators in TMEM; prec 1 = tf32, 3 = 3xTF32 split (fp32-grade).
 * w_hi / w_lo: float[K][nq][npad][4], nq = 4 * ceil(cin / 16): zero-padded to whole 16-channel slabs (see csrc/spconv_tc.cu). */
size_t ep_spconv_tc_workspace_bytes(int64_t m_out, int npad, int K);
int ep_spconv_tc_fwd(const float* in, int ld_in, int cin, const int32_t* nbr, int K, const float* w_hi,
                     const float* w_lo, int npad, int cout, const float* bias, float* out, int ld_out, int64_t m_out,
                     float* bn_partial, int prec, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* Shipped tensor-core variant (csrc/spconv_hl.cu): operands are PRE-SPLIT half pairs -- x = h + l * 2^-11, h = fp16(x), 22
 * significant bits like 3xTF32 at half the bytes -- stored in 32-channel slabs of 128 bytes [32 h | 32 l], the SWIZZLE_128B row
 * of a K-major tcgen05 operand; rows are gathered through the neighbour table by 16-byte cp.async straight into that layout
 * (default; EPRECON_HL_PRODUCER=tma selects the measured-slower cp.async.bulk.tensor tile::gather4 producer), weight slabs by
 * TMA, tcgen05.mma kind::f16 with fp32 accumulators in TMEM.  in_hl [m_in][nslab][64] halfs, nslab = ceil(cin / 32);
 * w_hl [K][nslab][npad][64] halfs; npad = cout rounded up to 16 (to 128 when larger). */
int ep_hl_slabs(int c);
int ep_hl_split_rows(const float* src, int ld_src, int c, int64_t m, uint16_t* dst, int32_t* overflow, cudaStream_t stream);
int ep_hl_affine_act(const float* a, int ld_a, const float* ss_a, const float* b, int ld_b, const float* ss_b, int relu, int64_t m,
                     int c, float* out, int ld_out, uint16_t* out_hl, int32_t* overflow, cudaStream_t stream);
int ep_hl_segment_mean(const float* feat, int ld_in, int c, const int32_t* perm, const int32_t* seg_start, const int32_t* seg_end,
                       int64_t m, float* out, int ld_out, uint16_t* out_hl, int32_t* overflow, cudaStream_t stream);
size_t ep_spconv_hl_workspace_bytes(int64_t m_out, int npad, int K);
int ep_spconv_hl_fwd(const uint16_t* in_hl, int64_t m_in, int cin, const int32_t* nbr, int K, const uint16_t* w_hl, int npad,
                     int cout, const float* bias, float* out, int ld_out, int64_t m_out, float* bn_partial, void* workspace,
                     size_t workspace_bytes, int neg_row_mode, cudaStream_t stream);
int ep_spconv_hl_fused_fwd(const uint16_t* in_hl, int64_t m_in, int cin, const int32_t* nbr, int K, const uint16_t* w_hl, int npad,
                           int cout, const float* bias, float* out, int ld_out, int64_t m_out, float* bn_partial, void* workspace,
                           size_t workspace_bytes, int neg_row_mode, int32_t* counters, int counters_len, const float* gamma,
                           const float* beta, float eps, float* ss_out, cudaStream_t stream);
int ep_spconv_hl_launches(int64_t m_out, int cin, int npad, int K, int have_counters, int want_ss);
int ep_hl_set_timeline(void* dev_buffer);   /* debug: clock64 timeline of one CTA of every following ep_spconv_hl launch */
int ep_hl_debug_code(void);   /* last failure site of ep_spconv_hl_fwd: 1 map A, 2 map B, 3 smem attribute, 4 launch */
int ep_hl_probe_gather4(const uint16_t* in_hl, int64_t m_in, int nslab, const int32_t* rows128, int slab, void* out16k,
                        int32_t* status, cudaStream_t stream);
int ep_colstats(const float* x, int ld, int64_t m, int c, float* bn_partial, cudaStream_t stream);
int ep_bn_finalize(const float* bn_partial, int num_row_tiles, int c, int64_t m, float eps, const float* gamma,
                   const float* beta, float* scale_shift, float* mean_var, cudaStream_t stream);

/* ---- row-wise / element-wise ----------------------------------------------------------------------------------
 * BatchNorm apply + residual + ReLU, LayerNorm variants, ConvGRU gates (models/modules.py:207-222), concats,
 * coordinate transforms (models/neucon_network.py:387-398, models/gru_fusion.py:331-337), x8 upsample (:193-214). */
int ep_affine_act(const float* a, int ld_a, const float* ss_a, const float* b, int ld_b, const float* ss_b, int relu,
                  int64_t m, int c, float* out, int ld_out, cudaStream_t stream);
int ep_layernorm(const float* x, int ld_x, const float* res, int ld_res, int relu_before, const float* gamma,
                 const float* beta, float eps, int relu_after, int64_t m, int c, float* out, int ld_out,
                 cudaStream_t stream);
/* train-mode BatchNorm2d of the dense 2-D fusion (models/modules.py:313-399; occupancy_initialization.py:41-58): batch
 * statistics over (N,H,W), optional ReLU + residual before (Conv2d_Residual_Block), optional ReLU after; 2 launches. */
size_t ep_bn2d_workspace_bytes(int c);
int ep_bn2d_train(const float* x, const float* res, int relu_pre, int n_img, int c, int hw, const float* gamma, const float* beta,
                  float eps, int relu_post, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int ep_gru_rh(const float* r_pre, int ld_r, const float* h, int ld_h, const float* x, int ld_x, int64_t m, int c,
              float* out, int ld_out, cudaStream_t stream);
int ep_gru_out(const float* z_pre, int ld_z, const float* q_pre, int ld_q, const float* h, int ld_h, int64_t m, int c,
               float* out, int ld_out, cudaStream_t stream);
int ep_gather_rows(const float* src, int ld_src, const int32_t* index, int shift, float fill, int64_t m, int c,
                   float* out, int ld_out, cudaStream_t stream);
int ep_gather_coords(const int32_t* src, const int32_t* index, int64_t m, int32_t* out, cudaStream_t stream);
int ep_aligned_coords(const int32_t* coords, int64_t n, const float* origin, float voxel_size, const float* w2ac,
                      int zero_batch, float* out, cudaStream_t stream);
int ep_upsample8(const int32_t* coords, int64_t n, int interval, int32_t* out, cudaStream_t stream);
int ep_threshold_flags(const float* x, int ld, int64_t n, float thr, int mode, uint8_t* flags, cudaStream_t stream);

/* ---- occupancy-initialisation pruning (models/neucon_network.py:264,298-318; erode/dilate :216-228) ---------- */
int ep_scatter_selected(const float* logit, int ld, const int32_t* src, int64_t n, float thr, uint8_t* fine,
                        cudaStream_t stream);
int ep_init_prune(const uint8_t* fine, int bs, int coarse_dim, int out_scale, int32_t* out_coords, int32_t* out_count,
                  cudaStream_t stream);

/* ---- GRU-fusion sparse union (models/gru_fusion.py:67-96,321-322; utils.py:176-180) -------------------------- */
int ep_fill_i32(int32_t* p, int64_t n, int32_t v, cudaStream_t stream);
int ep_scatter_rows_to_volume(const int32_t* coords, int coord_width, int coord_col0, int64_t n, int div, int ox,
                              int oy, int oz, int dx, int dy, int dz, const float* feat, int ld, int c, int mode,
                              int32_t* vol, uint8_t* valid, cudaStream_t stream);
int ep_union_flags(const int32_t* vol_a, const int32_t* vol_b, int64_t n, uint8_t* flags, cudaStream_t stream);
int ep_union_sites(const int32_t* sites, int64_t u, int dy, int dz, int batch, int scale, const int32_t* vol_a,
                   const int32_t* vol_b, int32_t* out_coords, int32_t* row_a, int32_t* row_b, cudaStream_t stream);

/* global-volume merge on the holder rank (configs[3]): in-order substitute-inside-bounding-volume rule of
 * GRUFusion(direct_substitute=True) (models/gru_fusion.py:93-94,198-204) over all gathered rows at once. */
int ep_merge_substitute_flags(const int32_t* rows, int64_t n, const int32_t* frag_start, const int32_t* boxes, int n_fragments,
                              uint8_t* flags, cudaStream_t stream);

/* ---- panoptic level alignment (models/neucon_network.py:516-544): hash-free parent marking instead of an O(N*M) compare */
int ep_mark_parents(const int32_t* coords, int64_t n, int step, int dx, int dy, int dz, int bs, uint8_t* vol,
                    cudaStream_t stream);
int ep_lookup_marks(const int32_t* coords, int64_t n, int step, int dx, int dy, int dz, int bs, const uint8_t* vol,
                    uint8_t* flags, cudaStream_t stream);

/* ---- panoptic decoder: masked cross-attention (models/mask3dformer.py:70-130,392-397 over nn.MultiheadAttention).
 * One pass over the keys: score -> mask -> online softmax -> weighted sum; per-chunk partials merged in fixed order.
 * q [n_queries, n_heads*head_dim] projected + unscaled; k, v [n_keys, ld_kv] projected; blocked uint8 [n_queries, n_keys]
 * (non-zero = the query may not attend to the key; NULL = no mask); out [n_queries, n_heads*head_dim].
 * Supported: head_dim == 6, n_queries <= 96, n_heads <= 8 (the reference decoder: 48 channels, 8 heads, 80 queries). */
size_t ep_masked_attention_workspace_bytes(int64_t n_keys, int n_heads);
int ep_masked_attention(const float* q, const float* k, const float* v, int ld_kv, const uint8_t* blocked, int64_t n_keys,
                        int n_queries, int n_heads, int head_dim, float scale, float* out, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream);
/* same, with a per-query flag (int32 [n_queries], optional): 0 = every key of that query is blocked, so the query ignores the
 * mask -- the reference's "fully blocked query attends everywhere" rewrite (models/mask3dformer.py:392) without a pass over
 * the flags */
int ep_masked_attention_flagged(const float* q, const float* k, const float* v, int ld_kv, const uint8_t* blocked,
                                const int32_t* row_unblocked, int64_t n_keys, int n_queries, int n_heads, int head_dim, float scale,
                                float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- scene TSDF -> mesh: GPU marching cubes + nearest-voxel labels, replaces the CPU tail of the reference's
 * mesh export (utils.py:231-247 skimage.measure.marching_cubes + np.round / np.clip lookup; volumes from gru_fusion.py:217-257).
 * vol f32 [dx,dy,dz]; edge_flags uint8 [3*n] (3 * voxel + axis: the surface crosses the grid edge leaving the voxel along
 * +axis -> one mesh vertex); cell_ntri uint8 [n]; vertices / faces are written at offsets from stable compactions / scans of
 * those, so their order is deterministic.  Case table: csrc/mc_table.cuh (derived by tools/gen_mc_table.py). */
int ep_mc_classify(const float* vol, int dx, int dy, int dz, float level, uint8_t* edge_flags, uint8_t* cell_ntri,
                   cudaStream_t stream);
int ep_mc_vertices(const float* vol, int dx, int dy, int dz, float level, const int32_t* edge_index, int64_t n_verts, float* verts,
                   float* normals, const int32_t* sem_vol, const int32_t* inst_vol, int32_t* sem_out, int32_t* inst_out,
                   cudaStream_t stream);
int ep_mc_faces(const float* vol, int dx, int dy, int dz, float level, const int32_t* cell_index, const int32_t* tri_offset,
                int64_t n_cells, const int32_t* edge_pos, int32_t* faces, cudaStream_t stream);

/* ---- panoptic decoder body as one native call (csrc/decoder.cu): MultiScaleMaskedTransformerDecoder.forward
 * (models/mask3dformer.py:337-445; layers :70-200; models/voxel_position_encoding.py:42-146) for the reference configuration
 * (48 channels, 8 heads, 80 queries, 6 layers, feed-forward 192, 20 classes).  desc: int64 parameter pointers, layout in
 * eprecon_b200/executor.py::_build_decoder.  pred_logits f32 [7][80][21] (prediction on query_feat + one per layer),
 * pred_masks f32 [80][n2] (last layer), aux_masks f32 [6][80][n2] or NULL. */
size_t ep_exec_decoder_workspace_bytes(int64_t n0, int64_t n1, int64_t n2);
int ep_exec_decoder(const int64_t* desc, const float* rows0, int ld0, const float* rows1, int ld1, const float* rows2, int ld2,
                    const int64_t* xyz0, const int64_t* xyz1, const int64_t* xyz2, int64_t n0, int64_t n1, int64_t n2,
                    const float* mask_rows, int ld_mask, const int64_t* index0, const int64_t* index1, float ex, float ey,
                    float ez, float* pred_logits, float* pred_masks, float* aux_masks, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream);

/* ---- small index helpers used by the native executor --------------------------------------------------------
 * ep_csr_expand: segments of a (voxel id)-keyed sort -> dense per-voxel [s0, s1) ranges (s0/s1 pre-zeroed; the scatter
 * the reference gets from torchsparse spcount/spvoxelize, ops/torchsparse_utils.py:51-58).  ep_translate_index:
 * out[i] = idx[i] >= 0 ? rank[idx[i]] : -1 (the ConvGRU.convr stale-cache view, models/modules.py:216-217). */
int ep_csr_expand(const uint64_t* keys_sorted, const int32_t* seg_start, const int32_t* seg_end, int64_t n_segments,
                  int32_t* s0, int32_t* s1, cudaStream_t stream);
int ep_translate_index(const int32_t* idx, int64_t n, const int32_t* rank, int32_t* out, cudaStream_t stream);

/* ---- native executor: one call per reference module (csrc/executor.cu) -----------------------------------------
 * Each function runs the complete launch program of one reference module forward() -- the same launchers, order and
 * arguments as the Python mirrors in eprecon_b200/modules.py, bit-identical results -- with every temporary carved
 * from the caller's scratch `arena` (256-byte aligned; EP_ERR_WORKSPACE if too small: grow and call again, the call
 * has no side effects besides `out`).  UNLIKE the per-kernel entry points they synchronise `stream` at the
 * data-dependent size read-backs (voxel counts).  `desc` / `globals`: flat int64 parameter descriptors, layout in
 * eprecon_b200/executor.py.  `stats` (optional, int64[2]) receives the arena high-water mark and the launch count.
 *   ep_exec_spvcnn     <- SPVCNN.forward               (models/modules.py:138-175)
 *   ep_exec_gru_level  <- the two ConvGRU.forward of a GRUFusion level (models/gru_fusion.py:339-349, modules.py:200-222)
 *   ep_exec_linear4x   <- Linear4xTrans.forward heads  (models/modules.py:298-311)
 *   ep_exec_init_head  <- Occupancy_Initialization.forward sparse head (models/occupancy_initialization.py:131-176) */
size_t ep_exec_launch_count(void);
int ep_exec_desc_check(int kind, const int64_t* desc);   /* host-only descriptor validation, no GPU needed */
/* per-launch CUDA-event timing of the sparse-conv family inside the executor (calling host thread only): enable(1), run
 * fragments, collect() -> number of records; meta int64 [cap,7] = {K, cin, cout, m_in, m_out, valid pairs, impl}, ms [cap]. */
int ep_exec_profile_enable(int on);
/* width tier of the executor's Z-order sort keys (0: 24 bits, 1: 32 bits, 2: 64 bits; raised automatically, sticky);
 * set < 0 only queries; returns the tier in force before the call */
int ep_exec_key_tier(int set);
int ep_exec_profile_collect(int64_t* meta, float* ms, int cap);
int ep_exec_spvcnn(const int64_t* desc, const int64_t* globals, const float* pts, const float* feat, int ld_feat,
                   int64_t n, float vres, float* out, int ld_out, void* arena, size_t arena_bytes, int64_t* stats,
                   cudaStream_t stream);
int ep_exec_gru_level(const int64_t* desc, const int64_t* globals, const float* pts, int64_t u, float vres,
                      const float* h, int ld_h, const float* x, int ld_x, float* out, int ld_out, void* arena,
                      size_t arena_bytes, int64_t* stats, cudaStream_t stream);
int ep_exec_linear4x(const int64_t* desc, const float* x, int ld_x, int64_t m, const int64_t* outs, void* arena,
                     size_t arena_bytes, int64_t* stats, cudaStream_t stream);
int ep_exec_init_head(const int64_t* desc, const int64_t* globals, const float* var, int ld_var, const int32_t* coords,
                      int64_t m, int sx, int sy, int sz, float* occ, void* arena, size_t arena_bytes, int64_t* stats,
                      cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EPRECON_B200_H_ */
