/* vdetr_b200.h -- C ABI of libvdetr_b200.so (sm_100a only).
 *
 * Drop-in boundary for the V-DETR Vertex-RPE decoder hot path.  The reference has no C-ABI plugin; its
 * boundary is (1) the pybind11 module `pointnet2._ext`
 * (third_party/pointnet2/_ext_src/src/bindings.cpp:9-21) and (2) the nn.Module classes of
 * models/vdetr_transformer.py.  Every entry point below names the reference interface it replaces.
 *
 * Conventions:
 *   - plain C types only; all pointers are DEVICE pointers unless named h_*; tensors are dense,
 *     row-major, in the layout written next to each argument;
 *   - the caller allocates every output and workspace; the library owns nothing across calls;
 *   - every function enqueues on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising; functions are re-entrant across streams;
 *   - return value: 0 = ok, VDETR_ERR_* (>0x1000) = bad argument, otherwise a cudaError_t value.
 *     The library never calls exit() (the reference does: include/cuda_utils.h:32-41).
 */
#ifndef VDETR_B200_H
#define VDETR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDETR_OK 0
#define VDETR_ERR_BAD_ARG 0x1001
#define VDETR_ERR_UNSUPPORTED 0x1002
#define VDETR_ERR_WORKSPACE 0x1003
#define VDETR_ERR_NO_DRIVER 0x1004

/* Library / build identification: returns e.g. "vdetr_b200 0.1 sm_100a". */
const char* vdetr_version(void);
/* Human readable text for a non-zero return code of this library. */
const char* vdetr_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * pointnet2 seed ops  (replace third_party/pointnet2/_ext_src)
 * ---------------------------------------------------------------------------------------------- */

/* furthest_point_sampling(points[B,N,3] f32, nsamples) -> int32[B,M]
 * Replaces: src/sampling.cpp:67-88 + src/sampling_gpu.cu:72-232 (bindings.cpp:12).
 * Bit-exact indices incl. the reference's tie-break and its |p|^2 <= 1e-3 skip rule.
 * workspace: vdetr_pn2_fps_workspace_bytes(B,N,M) bytes (may be 0); idx is fully overwritten. */
size_t vdetr_pn2_fps_workspace_bytes(int B, int N, int M);
int vdetr_pn2_fps(const float* xyz /*[B,N,3]*/, int B, int N, int M, int32_t* idx /*[B,M]*/,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Ragged batch: scene b owns rows [offsets[b], offsets[b+1]) of xyz [total_n,3] (offsets: B+1 int32 on the DEVICE, so no
 * host synchronisation is needed to call this).  Every scene is sampled exactly like a B = 1 call of vdetr_pn2_fps on its
 * own points; idx [B,M] holds scene-local indices.  Replaces the Python loop over scenes with one B = 1 launch each at
 * models/model_vdetr.py:282-316.  max_n = an upper bound of the per-scene point count known to the caller (it sizes the
 * thread-block clusters); a scene with more points gets idx = -1, an empty scene idx = 0.
 * workspace: vdetr_pn2_fps_ragged_workspace_bytes(B, total_n, max_n, M) (0 unless max_n exceeds the cluster kernel). */
size_t vdetr_pn2_fps_ragged_workspace_bytes(int B, int total_n, int max_n, int M);
int vdetr_pn2_fps_ragged(const float* xyz /*[total_n,3]*/, const int32_t* offsets /*[B+1]*/, int B, int total_n, int max_n, int M,
                         int32_t* idx /*[B,M]*/, void* workspace, size_t workspace_bytes, void* stream);

/* gather_points(points[B,C,N], idx[B,M]) -> out[B,C,M]        (src/sampling.cpp:17-40, sampling_gpu.cu:11-33) */
int vdetr_pn2_gather(const float* points, const int32_t* idx, int B, int C, int N, int M, float* out,
                     void* stream);
/* gather_points_grad(grad_out[B,C,M], idx[B,M], N) -> grad_points[B,C,N]  (src/sampling.cpp:42-66,
 * sampling_gpu.cu:37-60).  grad_points must be zero-filled by the caller (the reference does torch::zeros). */
int vdetr_pn2_gather_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int M,
                          float* grad_points, void* stream);

/* ball_query(new_xyz[B,M,3], xyz[B,N,3], radius, nsample) -> int32 idx[B,M,nsample]
 * Replaces: src/ball_query.cpp:11-35 + src/ball_query_gpu.cu:12-57.  idx is fully written (rows with no
 * neighbour are zero, like the reference's torch::zeros). Bit-exact. */
int vdetr_pn2_ball_query(const float* new_xyz, const float* xyz, int B, int N, int M, float radius,
                         int nsample, int32_t* idx, void* stream);

/* group_points(points[B,C,N], idx[B,M,S]) -> out[B,C,M,S]    (src/group_points.cpp:15-38, group_points_gpu.cu:11-42) */
int vdetr_pn2_group(const float* points, const int32_t* idx, int B, int C, int N, int M, int S, float* out,
                    void* stream);
/* group_points_grad(grad_out[B,C,M,S], idx[B,M,S], N) -> grad_points[B,C,N] (zero-filled by caller)
 * (src/group_points.cpp:40-63, group_points_gpu.cu:46-78) */
int vdetr_pn2_group_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int M, int S,
                         float* grad_points, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Vertex-RPE cross-attention core  (replaces the body of GlobalShareCrossAttention.forward,
 * models/vdetr_transformer.py:708-753, between the q/k/v Linear layers and `proj`)
 * ---------------------------------------------------------------------------------------------- */

typedef struct VdetrXattnShape {
  int B;        /* scenes */
  int nQ;       /* queries per scene */
  int nK;       /* key tokens per scene */
  int H;        /* query heads (must be 4) */
  int hd;       /* head dim (must be 64) */
  int grid_n;   /* table points per axis (10 for rpe_quant=bilinear_4_10) */
  float log_scale;  /* args.log_scale (512) */
  float max_value;  /* 4 for bilinear_4_10 */
  int rotate;   /* 1: angle_type=="object_coords": deltas are rotated by ref_angle per query */
  int kv_heads; /* 1 = shared K/V head (MQA, cross attention and ShareSelfAttention); H = per-head K/V (MHA) */
  int has_bias; /* 1 = add the Vertex-RPE bias; 0 = plain attention (decoder self-attention) */
} VdetrXattnShape;

/* Forward:  O = softmax_k(Q K^T + rpe(ref_pts, xyz, tables)) V        (fp32 in / fp32 out interface)
 *   q       [B,nQ,H,hd] f32, ALREADY multiplied by hd^-0.5            (vdetr_transformer.py:736-738); q, k, v (and dout in
 *           the backward) must be 16-byte aligned
 *   k, v    [B,nK,kv_heads,hd] f32                                    (:734-735)
 *   xyz     [B,nK,3] f32 ; ref_pts [B,nQ,8,3] f32 ; ref_angle [B,nQ] f32 or NULL   (:708-720)
 *   tables  [8,grid_n,grid_n,grid_n,H] f32 = cpb_mlps[i](relative_coords_table), a=z,b=y,c=x (:725)
 *   out     [B,nQ,H,hd] f32  (heads concatenated h-major = the input of `proj`, :755)
 *   lse     [B,H,nQ] f32 natural-log-sum-exp of the logits (saved for backward)
 *   bias_save  optional (may be NULL): vdetr_xattn_bias_save_bytes(&shape, impl) bytes in which the fused kernel
 *           leaves the fp32 bias of every (query,key) pair for the backward, which then streams it back instead
 *           of recomputing it (training: 16 B per pair of activation memory buys ~5x on the backward pass-1
 *           kernel; the reference keeps ~2.4 GB of autograd intermediates per layer-scene for the same purpose).
 *   dropout_p, dropout_seed: nn.Dropout on the attention probabilities (attn_drop, vdetr_transformer.py:751-752;
 *           main.py:75 trains with 0.1).  dropout_p in [0,1); when > 0, dropout_seed is a DEVICE pointer to one 64-bit
 *           seed (device memory so that the call can be captured in a CUDA graph and re-seeded between replays).
 *           The keep mask of element (row, key) is a pure function of (seed, row, key) -- Philox4x32-10,
 *           csrc/philox.cuh -- so the backward regenerates it from the same seed; kept probabilities are scaled by
 *           1/(1-p).  dropout_p = 0 (eval) ignores the seed.  impl 1 does not implement dropout.
 * impl: 0 = tcgen05/TMA fused kernel (product path), 1 = SIMT validation kernel.
 * workspace: vdetr_xattn_fwd_workspace_bytes(&shape, impl). */
size_t vdetr_xattn_fwd_workspace_bytes(const VdetrXattnShape* s, int impl);
size_t vdetr_xattn_bias_save_bytes(const VdetrXattnShape* s, int impl);   /* 0 when the option does not apply */
int vdetr_xattn_fwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v,
                    const float* xyz, const float* ref_pts, const float* ref_angle, const float* tables,
                    float* out, float* lse, float* bias_save, float dropout_p, const uint64_t* dropout_seed,
                    void* workspace, size_t workspace_bytes, int impl, void* stream);

/* Backward of the same op:
 *   in : q,k,v,xyz,ref_pts,ref_angle,tables as forward; out, lse from forward; dout [B,nQ,H,hd];
 *        bias_saved = the forward's bias_save buffer or NULL (NULL: the forward kernel is run once more into a
 *        transient buffer inside the workspace -- pass bias_is_saved = 0 to the workspace query);
 *        dropout_p / dropout_seed: the values the forward was called with
 *   out: dq [B,nQ,H,hd], dk, dv [B,nK,kv_heads,hd], dtables [8,n,n,n,H]  -- all fully overwritten.
 *   No gradient is produced for xyz / ref_pts (detached in the reference, vdetr_transformer.py:369-412). */
size_t vdetr_xattn_bwd_workspace_bytes(const VdetrXattnShape* s, int impl, int bias_is_saved);
int vdetr_xattn_bwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v,
                    const float* xyz, const float* ref_pts, const float* ref_angle, const float* tables,
                    const float* out, const float* lse, const float* dout, const float* bias_saved, float dropout_p,
                    const uint64_t* dropout_seed, float* dq, float* dk, float* dv, float* dtables, void* workspace,
                    size_t workspace_bytes, int impl, void* stream);

/* Bias only (debug / return_attn_weights path): rpe [B,H,nQ,nK] f32  (vdetr_transformer.py:708-731). */
int vdetr_rpe_bias(const VdetrXattnShape* s, const float* xyz, const float* ref_pts, const float* ref_angle,
                   const float* tables, float* rpe, void* stream);

/* Optional per-kernel timing for benchmarks: when enabled, CUDA events are recorded on the launch stream
 * around [0] the fused forward kernel, [1] the backward pass-1 kernel, [2] the dTables kernel, [3] the backward
 * pass-2 (dK / dV) kernel (cross-attention launches only; at most 512 launches each between reads).  vdetr_timing_read synchronises on the recorded events, returns the summed
 * device time in ms and the launch counts, and resets the counters. */
/* Number of kernels this library has launched (host-side count of its own launches; cuBLAS GEMMs are not included). */
unsigned long long vdetr_launch_count(int reset);
int vdetr_timing_enable(int enable);
int vdetr_timing_read(float* total_ms /*[4]*/, int* launches /*[4]*/);

/* Adjoint of vdetr_rpe_bias w.r.t. the tables: dtables[8,n,n,n,H] = sum_{b,q,k} dbias[b,q,k,h] * w_corner
 * (what autograd of F.grid_sample returns for its input at vdetr_transformer.py:727-731).
 * dbias is given pair-major: [B,nQ,nK,H] f32.  dtables is fully overwritten. */
size_t vdetr_rpe_dtables_workspace_bytes(const VdetrXattnShape* s);
int vdetr_rpe_dtables(const VdetrXattnShape* s, const float* xyz, const float* ref_pts, const float* ref_angle,
                      const float* dbias, float* dtables, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension of a dense [rows, cols] f32 matrix (the decoder's nn.LayerNorm(256) calls:
 * models/vdetr_transformer.py:463-466 norm1-3, :586-606 FFNLayer.norm, :129 TransformerDecoder.norm).
 * cols must satisfy vdetr_layernorm_supported (128, 256, 384 or 512).
 *   fwd: y [rows,cols], mean [rows], rstd [rows] (saved for backward)       bwd: dx [rows,cols], dgamma / dbeta [cols]
 *   (dgamma / dbeta are fully overwritten).
 * Column sums that cross CTAs (dgamma / dbeta here, the BatchNorm statistics and gradients, vdetr_colsum) are reduced
 * without float atomics: per-CTA partial sums in `workspace` (vdetr_reduce_workspace_floats(cols) floats, not
 * initialised by the caller) are added in a fixed order, so results are bit-reproducible run to run -- like the stock
 * nn.LayerNorm / nn.BatchNorm1d kernels the reference relies on. */
size_t vdetr_reduce_workspace_floats(int cols);
int vdetr_layernorm_supported(int cols);
int vdetr_layernorm_fwd(const float* x, const float* gamma, const float* beta, int rows, int cols, float eps, float* y,
                        float* mean, float* rstd, void* stream);
int vdetr_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int rows,
                        int cols, float* dx, float* dgamma, float* dbeta, float* workspace, void* stream);

/* out[c] = sum_r x[r, c] for a dense [rows, cols] f32 matrix (bias gradient of the token-major Linear layers).
 * workspace: vdetr_colsum_workspace_floats(cols) floats; deterministic (fixed summation order). */
size_t vdetr_colsum_workspace_floats(int cols);
int vdetr_colsum(const float* x, int rows, int cols, float* out, float* workspace, void* stream);

/* Training-mode BatchNorm1d + ReLU on token-major activations (the Conv1d(k=1)-BatchNorm1d-ReLU stacks of
 * models/helpers.py:17-33 and :74-141 evaluated per token).  `groups` independent BatchNorm layers of `cols` channels each
 * are normalised by one call (the 5 box heads of a decoder level, models/vdetr_transformer.py:256-300, evaluated
 * together): element (row r, group g, channel c) is x[g * group_stride + r * row_stride + c] (strides in floats, multiples
 * of 4) -- groups = 1, row_stride = cols is a plain dense [rows, cols] matrix; y / dy / dx use the same layout.
 * gamma, beta, mean, rstd, running_mean, running_var, dgamma, dbeta are [groups * cols].  Batch statistics over the rows;
 * running_mean / running_var (may be NULL) are updated in place with `momentum` and the unbiased variance, like
 * nn.BatchNorm1d.  y is the forward output (its sign is the ReLU mask).  cols: vdetr_bn_relu_supported (128, 256, 384,
 * 512).  workspace (both directions): vdetr_reduce_workspace_floats(groups * cols) floats.
 * dropout_p > 0 fuses the nn.Dropout that follows the ReLU in the reference's stacks (models/helpers.py:118-120,
 * --mlp_dropout): y = relu(bn(x)) * keep / (1 - p) with a Philox mask from the 64-bit seed at the DEVICE pointer
 * dropout_seed; the backward takes the same dropout_p and needs no mask (a dropped element is stored as 0). */
int vdetr_bn_relu_supported(int cols);
int vdetr_bn_relu_train_fwd(const float* x, const float* gamma, const float* beta, int rows, int cols, int groups,
                            long long group_stride, long long row_stride, float eps, float momentum, float dropout_p,
                            const uint64_t* dropout_seed, float* y, float* mean, float* rstd, float* running_mean,
                            float* running_var, float* workspace, void* stream);
int vdetr_bn_relu_train_bwd(const float* dy, const float* y, const float* x, const float* mean, const float* rstd,
                            const float* gamma, int rows, int cols, int groups, long long group_stride, long long row_stride,
                            float dropout_p, float* dx, float* dgamma, float* dbeta, float* workspace, void* stream);

/* Box decode of one decoder level (models/vdetr_transformer.py:244-333 for num_angle_bin = 1, angle == 0): head outputs
 * center_reg / size_reg [B,nQ,3] + the proposal boxes pre_*_normalized [B,nQ,3] + the scene extent dims_min / dims_max [B,3]
 * -> center, center_normalized, size, size_normalized, pre_center, pre_size [B,nQ,3], the 8 camera-frame corners
 * [B,nQ,8,3] (dataset_config.box_parametrization_to_corners, utils/box_util.py:294-358) and the lidar-frame corners
 * (convert_corners_camera2lidar, :98-102) that are the next layer's reference_point.  One kernel per direction instead of
 * ~35 elementwise launches.  Backward: gradients of center_reg / size_reg from the gradients of the five differentiable
 * outputs (any of them may be NULL = zero); the proposal boxes and the scene extent receive no gradient (detached in the
 * reference, :369-412). */
int vdetr_box_decode_fwd(const float* center_reg, const float* size_reg, const float* pre_center_normalized,
                         const float* pre_size_normalized, const float* dims_min, const float* dims_max, int B, int nQ,
                         float* center, float* center_normalized, float* size, float* size_normalized, float* pre_center,
                         float* pre_size, float* corners, float* ref_lidar, void* stream);
int vdetr_box_decode_bwd(const float* size, const float* pre_size, const float* dims_min, const float* dims_max,
                         const float* g_center, const float* g_center_normalized, const float* g_size,
                         const float* g_size_normalized, const float* g_corners, int B, int nQ, float* g_center_reg,
                         float* g_size_reg, void* stream);

/* Batched rectangular linear sum assignment (the Hungarian matching of criterion.py:205-228, which calls
 * scipy.optimize.linear_sum_assignment per scene on the host): cost [B,nQ,ngt] f32; nactual_gt [B] int32 on the DEVICE (NULL:
 * all ngt columns are real).  Scene b assigns every ground-truth column g < nactual_gt[b] a distinct query minimising the
 * total cost.  Outputs as the reference builds them: per_prop_gt_inds [B,nQ] int64 (0 where unmatched),
 * proposal_matched_mask [B,nQ] f32.  ngt <= nQ <= 4096, ngt <= 512.  Same algorithm and FP64 duals as scipy; exact ties
 * may resolve differently (same total cost). */
int vdetr_lsap(const float* cost, const int32_t* nactual_gt, int B, int nQ, int ngt, long long* per_prop_gt_inds,
               float* proposal_matched_mask, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Peer memory over NVLink (one process per GPU; csrc/peer.cu).  Replaces, for the data-parallel step of the reference,
 * DistributedDataParallel's NCCL gradient all-reduce + clip_grad_norm_ + AdamW (main.py:515-517, engine.py:105-108) and
 * nn.SyncBatchNorm's per-layer collectives (main.py:512-514).
 * Buffers are cudaMalloc allocations shared as 64-byte CUDA IPC handles; `*_ptrs` arguments are HOST arrays of `world`
 * device pointers, entry [rank] being this process's own allocation.  flags: VDETR_PEER_CHANNELS * VDETR_PEER_MAX_WORLD
 * uint32 per rank (zero-initialised by vdetr_peer_alloc); epoch: this rank's private device array of VDETR_PEER_CHANNELS
 * uint32 (zero-initialised).  Every rank must issue the same sequence of exchanging calls. */
#define VDETR_PEER_MAX_WORLD 8
#define VDETR_PEER_CHANNELS 4
int vdetr_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int vdetr_peer_open(const unsigned char* handle64, void** ptr);
int vdetr_peer_close(void* ptr);
int vdetr_peer_free(void* ptr);
/* 1 if a barrier ever timed out (20 s) waiting for a peer: the results of this process are then invalid */
int vdetr_peer_error(int* out);
int vdetr_peer_barrier(void* const* flag_ptrs, int rank, int world, uint32_t* epoch, int channel, void* stream);
/* bytes of the per-rank partial-norm exchange buffer of vdetr_adamw_flat_peer */
size_t vdetr_peer_norm_bytes(void);
/* Gradient exchange fused with the optimizer: rank `rank` owns elements [lo, hi) of the flat vectors (lo % 4 == 0, hi % 4 == 0
 * or hi == n; the shards of all ranks tile [0, n)).  Sums the gradients of the shard over all ranks (P2P loads, rank order),
 * computes the global gradient norm, scales by grad_scale_host (1 / world) and -- if max_norm > 0 -- by
 * min(1, max_norm / (norm + 1e-6)) like clip_grad_norm_, applies AdamW (see vdetr_adamw_flat) to the shard with the LOCAL
 * shard-sized moments m, v and scratch `reduced`, and stores the new parameters into every rank's parameter vector (P2P
 * stores).  Three flag barriers (channels 0-2) order it against the peers; norm_out [1] (optional) gets the norm. */
int vdetr_adamw_flat_peer(void* const* p_ptrs, void* const* g_ptrs, void* const* flag_ptrs, void* const* norm_ptrs, int rank, int world,
                          uint32_t* epoch, float* reduced, float* m, float* v, long long n, long long n_decay, long long lo,
                          long long hi, const float* lr, const float* step, float grad_scale_host, float max_norm, float* norm_out,
                          float beta1, float beta2, float eps, float weight_decay, void* stream);
/* SyncBatchNorm: floats of the per-rank statistics exchange buffer for up to `cap` channels (all groups of a launch together) */
size_t vdetr_peer_bn_slot_floats(int cap);
/* Merge BatchNorm statistics over all ranks (channel 3).  fwd: sum / sumsq are this rank's shifted sums of `groups` x `cols`
 * channels over `rows` rows (pivot = row 0 of x, group g at x + g * group_stride) and are rewritten in place so that they
 * describe the statistics of all ranks; bwd: gsum_g / gsum_b = sums of dgamma / dbeta over all ranks.  total_rows_out [1]. */
int vdetr_peer_bn_fwd(void* const* flag_ptrs, void* const* slot_ptrs, int rank, int world, uint32_t* epoch, int cap, float* sum,
                      float* sumsq, const float* x, int cols, int groups, long long group_stride, int rows, float* total_rows_out,
                      void* stream);
int vdetr_peer_bn_bwd(void* const* flag_ptrs, void* const* slot_ptrs, int rank, int world, uint32_t* epoch, int cap,
                      const float* dgamma, const float* dbeta, int cols_total, int rows, float* gsum_g, float* gsum_b,
                      float* total_rows_out, void* stream);
/* Library-wide SyncBatchNorm switch (the equivalent of nn.SyncBatchNorm.convert_sync_batchnorm, main.py:512-514): with a
 * context of world > 1, vdetr_bn_relu_train_fwd / _bwd exchange their statistics through vdetr_peer_bn_*; NULL switches off. */
typedef struct VdetrPeerCtx {
  void* flags[VDETR_PEER_MAX_WORLD];
  void* slots[VDETR_PEER_MAX_WORLD];
  int rank, world;
  uint32_t* epoch;
  int cap;
} VdetrPeerCtx;
int vdetr_bn_sync_set(const VdetrPeerCtx* ctx);

/* AdamW on flat buffers (replaces the multi-tensor torch.optim.AdamW the reference builds in optimizer.py:25 and steps in
 * engine.py:105-108): p, g, m, v are device arrays of n floats (16-byte aligned); elements [0, n_decay) are weight-decayed.
 * lr [1], step [1] (the 1-based step count as a float, already incremented) and the optional grad_scale [1] live in DEVICE
 * memory so that a captured CUDA graph follows a schedule; the gradient is multiplied by grad_scale_host * grad_scale[0]
 * (1 / world size, gradient-norm clipping of engine.py:105-106).  Same update as torch.optim.AdamW. */
int vdetr_adamw_flat(float* p, const float* g, float* m, float* v, long long n, long long n_decay, const float* lr,
                     const float* step, const float* grad_scale, float grad_scale_host, float beta1, float beta2, float eps,
                     float weight_decay, void* stream);

/* Developer aid: with VDETR_DT_CLOCKS=1 in the environment the dTables kernel sums the SM cycles each of its phases
 * takes ([0] records, [1] zero+B0, [2] histogram, [3] scan, [4] scatter, [5] accumulate) over all CTAs; this call
 * copies the 8 counters to the host and clears them (synchronises the device). */
int vdetr_debug_dt_clocks(unsigned long long* out8);

#ifdef __cplusplus
}
#endif
#endif /* VDETR_B200_H */
