/*
 * agnn.h -- C ABI of libagnn.so: the B200 (sm_100a) kernels behind AnalysisGNN's
 * heterogeneous message-passing hot path.
 *
 * The reference (manoskary/analysisgnn) is 100% Python and has no FFI: the seam
 * is a set of torch.nn.Module forward calls.  Each entry point below names the
 * reference code whose arithmetic it replaces (paths relative to the reference
 * root).  The host-side mirror of those modules lives in analysisgnn_b200/nn/ and
 * reaches this library through ctypes (analysisgnn_b200/_lib.py); INTEGRATION.md
 * shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless a parameter is
 *     documented as a host array (the small descriptor arrays are host arrays
 *     that are copied into kernel parameters at launch);
 *   - no allocation, no synchronisation, no host<->device copy inside: every
 *     call only enqueues kernels on `stream` (a cudaStream_t), so calls can be
 *     captured in CUDA graphs;
 *   - return value 0 = ok, negative = error (agnn_last_error() gives the text,
 *     thread-local);
 *   - feature matrices are row-major; `ld_*` are row strides in ELEMENTS;
 *     rows must be 16-byte aligned (ld * sizeof(elem) % 16 == 0) and the feature
 *     count a multiple of 4 (f32) / 8 (bf16);
 *   - index outputs are int32; COO inputs are int64 as PyTorch / PyG hand them over.
 */
#ifndef AGNN_H_
#define AGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* agnn_stream_t; /* cudaStream_t */

#define AGNN_OK 0
#define AGNN_ERR_ARG (-1)
#define AGNN_ERR_CUDA (-2)
#define AGNN_ERR_WORKSPACE (-3)
#define AGNN_ERR_UNSUPPORTED (-4)

#define AGNN_F32 0
#define AGNN_BF16 1

#define AGNN_MAX_SEG 32 /* COO segments per agnn_csr_build call */
#define AGNN_MAX_REL 16 /* relations fused in one gather / attention launch */

int agnn_version(void);
const char* agnn_last_error(void);

/* ------------------------------------------------------------------ CSR build
 * Replaces: the per-relation boolean mask + edge_index[:, mask] of
 * analysisgnn/models/core/hgnn.py:480-483, and the index_sort / CSC conversion
 * PyG's NeighborSampler does (third-party).
 *
 * One call converts up to AGNN_MAX_SEG COO segments to relation-major CSR.  For
 * segment s, rows are ordered by (relation, row) and ties keep INPUT ORDER
 * (stable), so results are bit-identical to a stable sort:
 *   rowptr[rowptr_off + r*(n_rows+1) + i] .. [.. + i + 1]  delimit, inside the
 *   segment's own block col[edge_off ..], the entries of row i of relation r;
 *   col[.]  = gathered-side node id, perm[.] = input edge position.
 * Edges whose etype is outside [0, n_rel) are dropped.  An edge with a node id
 * out of range sets *status to 1 and is dropped (status must be zeroed by the
 * caller; it is only ever written with 1).
 */
typedef struct agnn_coo {
  const int64_t* row;   /* reduce-side node id of each edge           */
  const int64_t* col;   /* gathered-side node id of each edge         */
  const int64_t* etype; /* relation id of each edge, or NULL (all 0)  */
  int64_t n_edges;
  int32_t n_rows; /* node count on the reduce side   */
  int32_t n_cols; /* node count on the gathered side */
  int32_t n_rel;
  int32_t reserved;
  int64_t rowptr_off; /* element offset of this segment's [n_rel*(n_rows+1)] block in rowptr */
  int64_t edge_off;   /* element offset of this segment's [n_edges] block in col / perm       */
  int64_t heavy_off;  /* element offset of this segment's [n_rel][2 * heavy_cap] block in `heavy` */
  int64_t count_off;  /* element offset of this segment's [n_rel] counters in `n_heavy`        */
  int64_t heavy_cap;  /* capacity per relation; n_edges / AGNN_HEAVY_ROW + 1 always suffices   */
} agnn_coo_t;

/* Rows with at least AGNN_HEAVY_ROW entries are also listed per (segment, relation) when `heavy` and `n_heavy` are
 * given: the relation's block of 2 * heavy_cap ints holds the row ids in ascending order, then (second half) the
 * exclusive prefix of their chunk counts ceil(deg / AGNN_HEAVY_CHUNK); `n_heavy` holds the count.
 * agnn_gather_reduce splits such rows across many warps, one chunk each, instead of letting one warp walk them alone
 * (hub nodes, Zipf-like degree tails); the prefix lets a warp find its chunk by binary search. */
#ifndef AGNN_HEAVY_ROW
#define AGNN_HEAVY_ROW 512
#endif
#ifndef AGNN_HEAVY_CHUNK
#define AGNN_HEAVY_CHUNK 256
#endif
/* the two constants this library was built with */
void agnn_heavy_params(int32_t* heavy_row, int32_t* heavy_chunk);

/* bytes of scratch agnn_csr_build needs for these segments (host arithmetic only) */
size_t agnn_csr_build_workspace(int n_seg, const agnn_coo_t* segs /* host */);

int agnn_csr_build(int n_seg, const agnn_coo_t* segs /* host */, int32_t* rowptr, int32_t* col,
                   int32_t* perm, int32_t* status, int32_t* heavy /* optional */, int32_t* n_heavy /* optional */,
                   void* workspace, size_t workspace_bytes, agnn_stream_t stream);

/* ------------------------------------------------------------ gather-reduce
 * Replaces: h[edge_index[1]] + torch_scatter.scatter(..., out=x.clone(),
 * reduce='mean') of SageConvScatter (analysisgnn/models/core/gnn.py:70-74), the
 * scatter_add pairs of MetricalConvLayer (gnn.py:511, 539) and MetricalGNN
 * (hgnn.py:406-407), onset pooling (analysisgnn/models/analysis.py:586), PyG
 * SAGEConv's mean aggregation (third-party) -- and, run on the transposed CSR,
 * the backward of each.  Warp-per-row segmented reduction, no atomics; the sum
 * over a row runs in CSR (= input edge) order, in fp32.
 *
 * For every row i and relation r:
 *     acc_r = sum_{k in rowptr_r[i] .. rowptr_r[i+1]}  w_k * src_r[col_r[k], :]
 *     w_k   = 1 / max(deg(col_r[k]), 1) from nbr_deg_rowptr if given, else 1
 * AGNN_COMBINE_CONCAT:  out[i, out_col_r : +F] = s_i * (self_add[i] + acc_r)
 * AGNN_COMBINE_SUM:     out[i, out_col_0 : +F] = self_add[i] + sum_r s_i,r * acc_r
 *     s = 1 / max(deg_r(i), 1) with AGNN_SCALE_MEAN, 1 with AGNN_SCALE_NONE
 * A relation flagged AGNN_REL_IDENTITY_IF_EMPTY that has no edges at all (checked on
 * the device: rowptr_r[n_rows] == rowptr_r[0], so no host sync is needed)
 * contributes src_r[i, :] itself, unscaled, without the self term: the reference's
 * E == 0 branch, z = W [x || h] (gnn.py:67-69).
 * If `copy` is given, out[i, copy_col : +F] = copy[i, :] (builds [x || S] rows).
 * If `out_lo` is given (fp32 only; same shape and stride as `out`), every value y is stored as the TF32
 * pair out = rna_tf32(y), out_lo = rna_tf32(y - out): the operand form of agnn_gemm's 3xTF32 mode,
 * produced here for free instead of by a separate agnn_split_tf32 pass over the matrix.
 */
typedef struct agnn_rel {
  const int32_t* rowptr; /* [n_rows + 1]                                        */
  const int32_t* col;    /* base the rowptr values index into                   */
  const void* src;       /* gathered matrix (column slice already applied)      */
  int64_t ld_src;
  const int32_t* nbr_deg_rowptr; /* optional [n_src + 1]: neighbour weights     */
  int32_t out_col;
  int32_t flags;
  const int32_t* heavy_rows; /* optional: this relation's block of `heavy` (agnn_csr_build): [heavy_cap] ascending row
                              * ids, then [heavy_cap] chunk prefix                                                   */
  const int32_t* n_heavy;    /* device counter that goes with heavy_rows                        */
  int64_t heavy_cap;         /* rows the block can list (it is 2 * heavy_cap ints long)         */
} agnn_rel_t;

#define AGNN_REL_IDENTITY_IF_EMPTY 1
/* hint for a launch of ONE relation: edges / rows < 1.5, so rows are reduced four at a time per warp (four source rows in
 * flight instead of min(degree, 4)); results are identical, any degree is still handled */
#define AGNN_REL_LOW_DEGREE 2

#define AGNN_SCALE_NONE 0
#define AGNN_SCALE_MEAN 1
#define AGNN_COMBINE_CONCAT 0
#define AGNN_COMBINE_SUM 1

int agnn_gather_reduce(int32_t n_rows, int32_t n_feat, int dtype, int scale, int combine, int n_rel,
                       const agnn_rel_t* rels /* host */, const void* self_add, int64_t ld_self,
                       const void* copy, int64_t ld_copy, int32_t copy_col, void* out, int64_t ld_out,
                       void* out_lo /* optional */, void* heavy_workspace /* optional */, size_t heavy_workspace_bytes,
                       agnn_stream_t stream);
/* Same launch with the concatenated output written as the fp16 hi / lo operand pair of agnn_gemm's F16X3 mode: out and
 * out_lo are fp16 matrices (ld_out in fp16 elements, a multiple of 8), every value is scaled by the power of two
 * derived from *pair_amax (a device scalar >= max |y| / 4 -- e.g. the amax of the gathered features: a mean of rows
 * cannot exceed it).  fp32 inputs, AGNN_COMBINE_CONCAT only; pair_amax == NULL is agnn_gather_reduce. */
int agnn_gather_reduce_f16(int32_t n_rows, int32_t n_feat, int dtype, int scale, int combine, int n_rel,
                           const agnn_rel_t* rels /* host */, const void* self_add, int64_t ld_self, const void* copy,
                           int64_t ld_copy, int32_t copy_col, void* out, int64_t ld_out, void* out_lo,
                           const float* pair_amax, void* heavy_workspace /* optional */, size_t heavy_workspace_bytes,
                           agnn_stream_t stream);
/* Same launch; amax_out (optional device scalar, zero it first) receives max(*amax_out, max |value written|) -- the
 * fp16 operand scale of the result for a later agnn_gemm without another pass over it (the gradient a backward gather
 * hands to the projection below it).  Plain fp32 / bf16 output only (out_lo == pair_amax == NULL). */
int agnn_gather_reduce_amax(int32_t n_rows, int32_t n_feat, int dtype, int scale, int combine, int n_rel,
                            const agnn_rel_t* rels /* host */, const void* self_add, int64_t ld_self, const void* copy,
                            int64_t ld_copy, int32_t copy_col, void* out, int64_t ld_out, void* out_lo,
                            const float* pair_amax, float* amax_out, void* heavy_workspace /* optional */,
                            size_t heavy_workspace_bytes, agnn_stream_t stream);
/* Scratch for the heavy-row path of a launch whose relations hold `total_edges` edges and `total_heavy_cap` heavy
 * slots in all: (total_edges / AGNN_HEAVY_CHUNK + total_heavy_cap) partial rows of n_feat floats.  Relations are
 * only split when heavy_rows is set AND a workspace is given; otherwise every row is walked by its own warp. */
size_t agnn_gather_heavy_workspace(int64_t total_edges, int64_t total_heavy_cap, int32_t n_feat);

/* Fused per-edge message + segmented reduction for the reference's alternative conv blocks
 * (analysisgnn/models/core/gnn.py): per row i of a single-relation CSR (rowptr / col as agnn_csr_build writes them),
 *   out0[i] = s_i * (self_add[i] + sum_{k in row i} msg(i, col[k])),  s_i = 1 / max(deg_i, 1) if mean else 1,
 * every message formed in registers -- no [E, F] tensor exists on either pass.  row* operands are indexed by the row,
 * nbr* operands by col[k]; fp32, n_feat a multiple of 4 up to 512.
 *   AGNN_EDGE_ABSDIFF       msg = |row0_i - nbr0_j|                       OnsetEmbedding (gnn.py:300-311), RelEdgeConv (:99-106)
 *   AGNN_EDGE_ABSDIFF_DROW  msg = row0_i * sign(row1_i - nbr0_j)          its gradient wrt the row operand (row0 = g, row1 = a)
 *   AGNN_EDGE_ABSDIFF_DNBR  msg = -nbr0_i * sign(nbr1_i - row0_j)         ... wrt the neighbour operand, on the TRANSPOSED CSR
 *   AGNN_EDGE_GATE          msg = sigmoid(row0_i + nbr0_j) * nbr1_j       ResGatedGraphConv (gnn.py:243-258)
 *   AGNN_EDGE_GATE_DROW     msg = row0_i * nbr1_j * s (1 - s), s = sigmoid(row1_i + nbr0_j)     (row0 = g, row1 = a)
 *   AGNN_EDGE_GATE_DNBR     transposed CSR, row0 = b_j, row1 = c_j, nbr0 = g_i, nbr1 = a_i: out0 = d b_j, out1 = d c_j */
#define AGNN_EDGE_ABSDIFF 0
#define AGNN_EDGE_ABSDIFF_DROW 1
#define AGNN_EDGE_ABSDIFF_DNBR 2
#define AGNN_EDGE_GATE 3
#define AGNN_EDGE_GATE_DROW 4
#define AGNN_EDGE_GATE_DNBR 5
int agnn_edge_op(int op, int32_t n_rows, int32_t n_feat, const int32_t* rowptr, const int32_t* col, const float* row0,
                 int64_t ld_row0, const float* row1, int64_t ld_row1, const float* nbr0, int64_t ld_nbr0,
                 const float* nbr1, int64_t ld_nbr1, const float* self_add, int64_t ld_self, int mean, float* out0,
                 int64_t ld_out0, float* out1, int64_t ld_out1, agnn_stream_t stream);

/* out[i, :] = base[i, :] (if given) + sum_r in[i, in_col_r : +F] / max(deg_r(i), 1)
 * -- the gradient of the self term of the mean_self reduction (gnn.py:74 backward).
 * Relations flagged AGNN_REL_IDENTITY_IF_EMPTY that are empty are skipped.  Uses rels[r].rowptr,
 * .out_col (as the column of `in`) and .flags only. */
int agnn_rowscale_sum(int32_t n_rows, int32_t n_feat, int dtype, int n_rel, const agnn_rel_t* rels /* host */,
                      const void* in, int64_t ld_in, const void* base, int64_t ld_base, void* out,
                      int64_t ld_out, agnn_stream_t stream);

/* ------------------------------------------------------------ HGT attention
 * Replaces, per destination node type, the edge-level pipeline of PyG's HGTConv
 * that graphmuse's HybridHGT runs (constructed at analysisgnn/models/analysis.py:
 * 444-453; third-party arithmetic): q[dst] . k[src] scores,
 * the segment softmax over every incoming edge of a destination (all relations
 * together) and the alpha-weighted sum of v[src].  One warp per destination row,
 * online softmax in registers, no per-edge tensors, no atomics.
 *
 *   s_e   = (q_i,h . k_e,h) * pscale[r(e), h]          (pscale = p_rel / sqrt(D), device array [n_rel*heads])
 *   out_i,h = sum_e exp(s_e - max_i,h) v_e,h / (sum_e exp(s_e - max_i,h) + 1e-16)
 *   row_max / row_den [n_dst*heads] keep max_i,h (in log2 units: score * log2 e, the kernels' softmax runs in base 2)
 *   and the denominator for the backward passes; opaque to callers.
 * k / v are the relation-specific keys / values of the SOURCE type, [n_src, heads*head_dim].
 */
#define AGNN_HGT_MAX_HEADS 16

typedef struct agnn_hgt_rel {
  const int32_t* rowptr;   /* [n_dst + 1]  CSR keyed on the destination (fwd, bwd_dst)          */
  const int32_t* col;      /* source ids                                                        */
  const int32_t* t_rowptr; /* [n_src + 1]  CSR keyed on the source (bwd_src)                    */
  const int32_t* t_col;    /* destination ids                                                   */
  const void* k;
  const void* v;
  int64_t ld_kv;
  void* dk; /* bwd_src outputs [n_src, heads*head_dim] */
  void* dv;
  int64_t ld_dkv;
  int32_t n_src;
  int32_t reserved;
} agnn_hgt_rel_t;

int agnn_hgt_attn_fwd(int32_t n_dst, int heads, int head_dim, int dtype, int n_rel,
                      const agnn_hgt_rel_t* rels /* host */, const void* q, int64_t ld_q, const float* pscale,
                      void* out, int64_t ld_out, float* row_max, float* row_den, agnn_stream_t stream);

/* Backward, destination side: delta_i,h = dout_i,h . out_i,h; dq; and per-block partial sums of
 * d pscale, dpscale_partial[agnn_hgt_attn_bwd_dst_blocks(n_dst)][n_rel*heads] (summed by the caller,
 * which keeps the reduction order fixed). */
int agnn_hgt_attn_bwd_dst_blocks(int32_t n_dst);
int agnn_hgt_attn_bwd_dst(int32_t n_dst, int heads, int head_dim, int dtype, int n_rel,
                          const agnn_hgt_rel_t* rels /* host */, const void* q, int64_t ld_q, const float* pscale,
                          const void* out, int64_t ld_out, const void* dout, int64_t ld_dout,
                          const float* row_max, const float* row_den, float* delta, void* dq, int64_t ld_dq,
                          float* dpscale_partial, agnn_stream_t stream);

/* Backward, source side: for every relation r and source row j, dk_r[j], dv_r[j] (written, not
 * accumulated) from the transposed CSR; alpha is recomputed from row_max / row_den. */
int agnn_hgt_attn_bwd_src(int heads, int head_dim, int dtype, int n_rel, const agnn_hgt_rel_t* rels /* host */,
                          const void* q, int64_t ld_q, const float* pscale, const void* dout, int64_t ld_dout,
                          const float* row_max, const float* row_den, const float* delta, agnn_stream_t stream);

/* ------------------------------------------------------------ optimizer step
 * Replaces, for data-parallel training, what Lightning runs after the reference's
 * training_step (analysisgnn/train/train_analysisgnn.py:246-255, gradient_clip_val=1.0
 * = clip_grad_norm_; AdamW from analysisgnn/models/analysis.py:1380): one pass for the
 * global gradient norm and one for  g <- g * grad_scale (1/world after the NCCL sum);
 * g <- g * min(1, max_norm / (||g|| + 1e-6));  AdamW (torch.optim.AdamW arithmetic).
 * Gradients and both moments live in flat fp32 arenas; parameters stay in their own
 * tensors and are reached through a device-resident chunk table (one block per chunk,
 * at most agnn_optim_chunk_elems() elements each; arena offsets multiples of 4).
 */
typedef struct agnn_param_chunk {
  void* param;          /* base of the parameter tensor (fp32, contiguous)        */
  int64_t param_off;    /* first element of this chunk inside the parameter       */
  int64_t arena_off;    /* first element of this chunk inside grad / m / v arenas */
  int32_t count;        /* elements in this chunk                                 */
  int32_t param_aligned; /* 1 if (param + param_off) is 16-byte aligned           */
} agnn_param_chunk_t;

int agnn_optim_chunk_elems(void);
int agnn_sumsq_blocks(int64_t n);
/* partials[agnn_sumsq_blocks(n)] = per-block sums of grad^2 (n multiple of 4, arena 16-byte aligned);
 * step_counter (optional, device int) is incremented by one -- the device-side step number that makes
 * the optimizer step replayable inside a CUDA graph. */
int agnn_sumsq_partials(const float* grad, int64_t n, float* partials, int* step_counter, agnn_stream_t stream);
/* chunks: DEVICE array [n_chunks]; the step number (from 1) is *step_dev if step_dev is given, else `step`; the
 * learning rate is *lr_dev if lr_dev is given (a device scalar the host's LR scheduler writes: the reference's
 * warm-up + cosine schedule, analysisgnn/models/analysis.py:1380-1400, keeps working under CUDA-graph replay), else `lr`;
 * max_norm <= 0 disables clipping; norm_out (optional, device) receives the norm of the averaged gradient. */
int agnn_adamw_clip_step(const agnn_param_chunk_t* chunks /* device */, int n_chunks, const float* grad, float* m,
                         float* v, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                         const int* step_dev, const float* lr_dev, float grad_scale, float max_norm,
                         const float* partials, int n_partials, float* norm_out, agnn_stream_t stream);

/* ------------------------------------------------------------ dense projections
 * Replaces the nn.Linear calls of the path and the GEMMs of their backward:
 * SageConvScatter.neigh_linear / .linear (analysisgnn/models/core/gnn.py:65, 75), project_dict
 * (analysisgnn/models/analysis.py:429-443, 575), PyG SAGEConv lin_l / lin_r and HGTConv kqv / out
 * projections (third party).  tcgen05.mma, TMEM accumulators, TMA operands (gemm.cu).
 *
 *   C[M,N] (+)= A[M,K] * B[K,N] + bias[N]   (optional ReLU)
 *
 * Layouts: AGNN_LAYOUT_K_MAJOR  = the operand is stored [M or N, K] row-major (K contiguous);
 *          AGNN_LAYOUT_MN_MAJOR = the operand is stored [K, M or N] row-major.
 *   forward      Y = X W^T          A = X  K-major,  B = W [out,in]  K-major
 *   grad input   dX = dY W          A = dY K-major,  B = W [out,in]  MN-major
 *   grad weight  dW = dY^T X        A = dY MN-major, B = X           MN-major  (use split_k)
 * Precision: AGNN_GEMM_TF32X3 = fp32 parity mode, operands given as TF32-exact hi / lo parts
 * (agnn_split_tf32), three tensor-core products accumulated in fp32; AGNN_GEMM_TF32 = hi only;
 * AGNN_GEMM_BF16 = bf16 operands (lo ignored), fp32 accumulation, fp32 or bf16 (AGNN_GEMM_OUT_BF16) output.
 * Row strides (lda, ldb, in elements) must be multiples of 16 bytes.  split_k > 1 needs
 * agnn_gemm_workspace() bytes; partial sums are reduced in a fixed order (deterministic).
 */
#define AGNN_GEMM_TF32X3 0
#define AGNN_GEMM_TF32 1
#define AGNN_GEMM_BF16 2
#define AGNN_GEMM_F16X3 3 /* fp32 parity on the f16 MMA: fp16 hi / lo operands with one power-of-two scale per tensor */
#define AGNN_LAYOUT_K_MAJOR 0
#define AGNN_LAYOUT_MN_MAJOR 1
#define AGNN_GEMM_RELU 1
#define AGNN_GEMM_ACCUMULATE 2
#define AGNN_GEMM_OUT_BF16 4

/* hi = rna_tf32(x), lo = rna_tf32(x - hi), both stored as fp32 bit patterns */
int agnn_split_tf32(const float* x, int64_t rows, int64_t cols, int64_t ld_x, float* hi, float* lo, int64_t ld_out,
                    agnn_stream_t stream);
/* suggested split_k for a problem (host arithmetic) and the workspace a given split_k needs */
int agnn_gemm_split_k(int precision, int64_t M, int64_t N, int64_t K);
size_t agnn_gemm_workspace(int precision, int64_t M, int64_t N, int64_t K, int split_k);
int agnn_gemm(int precision, int a_layout, int b_layout, int64_t M, int64_t N, int64_t K, const void* a_hi,
              const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb, void* c, int64_t ldc,
              const float* bias, int flags, int split_k, void* workspace, size_t workspace_bytes,
              agnn_stream_t stream);
/* AGNN_GEMM_F16X3 (same nn.Linear sites as agnn_gemm; twice the MMA rate of TF32X3, half the operand bytes):
 * operands are fp16 hi / lo pairs of s_a A and s_b B (lda / ldb in fp16 elements), where s = 2^k is derived from the
 * tensor's amax so that max |s x| lies in [2^13, 2^14); amax_a / amax_b point at those device scalars and the
 * epilogue divides by s_a s_b (exact) before the bias.  fp16 carries TF32's 11-bit significand, so
 * hi*hi + hi*lo + lo*hi has TF32X3's accuracy for elements down to 2^-17 of the tensor's amax (absolute error
 * 2^-25 of the scaled range below that).
 *   agnn_amax       *amax = max(*amax, max |x|)  (zero it first; several tensors may share one scalar)
 *   agnn_split_f16  hi = fp16(s x), lo = fp16(s x - hi), s from *amax */
int agnn_amax(const float* x, int64_t rows, int64_t cols, int64_t ld_x, float* amax, agnn_stream_t stream);
int agnn_split_f16(const float* x, int64_t rows, int64_t cols, int64_t ld_x, const float* amax, void* hi, void* lo,
                   int64_t ld_out, agnn_stream_t stream);
/* agnn_amax + agnn_split_f16 for up to AGNN_SPLIT_MULTI_MAX SMALL matrices (the weights of a grouped launch) in ONE
 * launch: a cluster of 8 CTAs per matrix exchanges its partial maxima through distributed shared memory.  *amax is
 * overwritten with max |x| of that matrix (it need not be zeroed). */
#define AGNN_SPLIT_MULTI_MAX 24
typedef struct agnn_split_item {
  const float* x; int64_t rows, cols, ld_x;
  void* hi; void* lo; int64_t ld_out; /* fp16 [rows, cols] each */
  float* amax;
} agnn_split_item_t;
int agnn_split_f16_multi(int n, const agnn_split_item_t* items /* host */, agnn_stream_t stream);
int agnn_gemm_scaled(int precision, int a_layout, int b_layout, int64_t M, int64_t N, int64_t K, const void* a_hi,
                     const void* a_lo, int64_t lda, const float* amax_a, const void* b_hi, const void* b_lo, int64_t ldb,
                     const float* amax_b, void* c, int64_t ldc, const float* bias, int flags, int split_k, void* workspace,
                     size_t workspace_bytes, agnn_stream_t stream);

/* Grouped launch: up to AGNN_GEMM_MAX_GROUP independent problems of ONE precision / layout combination in one
 * persistent launch -- the per-node-type projections (project_dict, analysis.py:429-443; PyG HeteroDictLinear in
 * HGTConv), the per-task heads (clf_dict, analysis.py:486-496, 546-569), the destination types of a message-passing
 * layer, the directions of a GRU layer, and the matching grad-input / grad-weight products of their backward.  Every
 * problem has its own sizes, operands, bias, flags and split_k.  Empty problems (M or N == 0) are skipped.
 *   workspace   split_k > 1: agnn_gemm_workspace() bytes for the partial tiles;
 *   tickets     optional device int32 array, ZERO on entry (and zero again on exit): with it the partials of a tile are
 *               added in split order by the CTA that stores the tile's last partial -- inside the same launch,
 *               deterministic; without it (NULL) a second kernel reduces them.  agnn_gemm_tickets() counters per
 *               problem, handed out in problem order.  One array must not be used by two launches that may overlap.
 *   amax_out    optional device scalar: *amax_out = max(*amax_out, max |C|) (zero it first) -- the F16X3 operand
 *               scale of C for a later product, without another pass over C.  Not with AGNN_GEMM_ACCUMULATE. */
#define AGNN_GEMM_MAX_GROUP 12
typedef struct agnn_gemm_problem {
  int64_t M, N, K;
  const void* a_hi; const void* a_lo; int64_t lda; const float* amax_a; /* amax_*: F16X3 only, else NULL */
  const void* b_hi; const void* b_lo; int64_t ldb; const float* amax_b;
  void* c; int64_t ldc;
  const float* bias;     /* optional [N] */
  int32_t flags;         /* AGNN_GEMM_RELU | AGNN_GEMM_ACCUMULATE | AGNN_GEMM_OUT_BF16 */
  int32_t split_k;       /* >= 1 */
  void* workspace; size_t workspace_bytes;
  float* amax_out;       /* optional */
} agnn_gemm_problem_t;
int64_t agnn_gemm_tickets(int64_t M, int64_t N, int split_k);
/* split counts for the problems of one grouped launch (host arithmetic): the group as a whole fills the machine */
int agnn_gemm_group_split_k(int precision, int n_problems, const int64_t* M, const int64_t* N, const int64_t* K,
                            int32_t* split_out);
int agnn_gemm_grouped(int precision, int a_layout, int b_layout, int n_problems, const agnn_gemm_problem_t* problems,
                      int32_t* tickets, int64_t n_tickets, agnn_stream_t stream);

/* The large products of a step (the fused message-passing layers' forward and grad-input GEMMs: M = nodes of a batch,
 * N >= 256) on CTA PAIRS: tcgen05.mma.cta_group::2, a 256 x 256 tile per pair, each CTA staging its 128 rows of A and
 * its half of B -- half the L2 -> shared-memory traffic per flop of agnn_gemm's 128 x 128 tiles, which those shapes
 * are bound by.  AGNN_GEMM_F16X3 operands only (fp16 hi / lo pairs + the two amax scalars), A K-major, B K-major or
 * MN-major, fp32 C with 16-byte aligned rows, optional bias / ReLU / amax_out; no split-K, no accumulate.
 * agnn_gemm_pair_supported tells whether a problem qualifies (host arithmetic). */
int agnn_gemm_pair_supported(int precision, int a_layout, int64_t M, int64_t N, int64_t K, int flags);
int agnn_gemm_pair(int b_layout, int64_t M, int64_t N, int64_t K, const void* a_hi, const void* a_lo, int64_t lda,
                   const float* amax_a, const void* b_hi, const void* b_lo, int64_t ldb, const float* amax_b, float* c,
                   int64_t ldc, const float* bias, int flags, float* amax_out, agnn_stream_t stream);

/* ------------------------------------------------------------ row-wise normalisation
 * agnn_layernorm_*: nn.LayerNorm of project_dict / project_enc (analysisgnn/models/analysis.py:429-443,
 * 474-485) and of the sequence branch (analysisgnn/models/cadence.py:249-260).  fp32, warp per row,
 * cols a multiple of 4 up to 1024.  Backward writes dx and per-block partial sums of d gamma / d beta,
 * [agnn_row_blocks(rows)][cols] each, which the caller adds up (fixed order).
 * agnn_l2norm_relu_*: F.normalize(p=2, eps) combined with the ReLU around it in MetricalGNN
 * (analysisgnn/models/core/hgnn.py:415, 421-422, 431): relu_first = normalize(relu(x)), else relu(normalize(x)).
 * agnn_colsum_partials: bias gradients, partials[agnn_row_blocks(rows)][cols].
 * When the optional final outputs are given (dgamma / dbeta / out) the partials are also summed, in block
 * order, by a second launch inside the call (dbeta_partials must directly follow dgamma_partials).
 */
int agnn_row_blocks(int64_t rows);
int agnn_layernorm_fwd(const float* x, int64_t ld_x, const float* gamma, const float* beta, float* y, int64_t ld_y,
                       float* mean, float* rstd, int64_t rows, int cols, float eps, agnn_stream_t stream);
int agnn_layernorm_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const float* gamma,
                       const float* mean, const float* rstd, float* dx, int64_t ld_dx, float* dgamma_partials,
                       float* dbeta_partials, float* dgamma /* optional [cols] */, float* dbeta /* = dgamma + cols */,
                       int64_t rows, int cols, agnn_stream_t stream);
/* LayerNorm feeding a projection (project_dict / project_enc / clf_dict / the sequence MLP are all
 * Linear -> ReLU -> LayerNorm -> Dropout -> Linear chains, analysisgnn/models/analysis.py:429-443, 474-496): the
 * (dropped-out) output is written directly as the fp16 hi / lo operand pair of the projection behind the norm
 * (pair_hi / pair_lo, ld_pair in fp16 elements, cols % 8 == 0) -- no fp32 output (y may be NULL), no dropout kernel,
 * no amax pass, no split pass.  The pair's scale comes from the bound |y| <= max|gamma| sqrt(cols - 1) + max|beta|
 * (times 1 / (1 - p)), which the kernel writes to *pair_amax for agnn_gemm_scaled.
 * Dropout (nn.Dropout in train mode; p = 0 or rng_state == NULL: none) is a counter-based mask: keep(element) is a
 * function of rng_state = {seed, step} (DEVICE uint64[2]; agnn_dropout_advance increments step, so a CUDA-graph replay
 * draws new masks), rng_stream (distinguishes call sites) and the element index; the backward recomputes it.
 * agnn_layernorm_bwd_dropout: dy is first multiplied by the same mask / (1 - p); amax_out (optional) receives max |dx|.
 * agnn_dropout_apply: y = keep ? x / (1 - p) : 0 for a plain matrix (both directions of the inter-layer dropout of
 * the sequence GRU, analysisgnn/models/cadence.py:249-251), optional amax_out.
 * agnn_split_f16_dropout: agnn_split_f16 of the dropped-out matrix in one pass (amax must bound |x| / (1 - p)). */
int agnn_layernorm_fwd_pair(const float* x, int64_t ld_x, const float* gamma, const float* beta, float* y /* optional */,
                            int64_t ld_y, float* mean, float* rstd, int64_t rows, int cols, float eps, void* pair_hi,
                            void* pair_lo, int64_t ld_pair, float* pair_amax, float dropout_p, const uint64_t* rng_state,
                            uint32_t rng_stream, agnn_stream_t stream);
int agnn_layernorm_bwd_dropout(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const float* gamma,
                               const float* mean, const float* rstd, float* dx, int64_t ld_dx, float* dgamma_partials,
                               float* dbeta_partials, float* dgamma, float* dbeta, int64_t rows, int cols,
                               float dropout_p, const uint64_t* rng_state, uint32_t rng_stream, float* amax_out,
                               agnn_stream_t stream);
int agnn_dropout_advance(uint64_t* rng_state, agnn_stream_t stream);
int agnn_dropout_apply(const float* x, int64_t ld_x, float* y, int64_t ld_y, int64_t rows, int cols, float dropout_p,
                       const uint64_t* rng_state, uint32_t rng_stream, float* amax_out, agnn_stream_t stream);
int agnn_split_f16_dropout(const float* x, int64_t rows, int64_t cols, int64_t ld_x, const float* amax, void* hi, void* lo,
                           int64_t ld_out, float dropout_p, const uint64_t* rng_state, uint32_t rng_stream,
                           agnn_stream_t stream);
/* agnn_split_f16_shifted: the pair of the matrix whose row r is input row r - shift of the same sequence (rows come in
 * sequences of seq_len; zeros where that row does not exist): the h_{t-1} / h_{t+1} operand of a GRU's recurrent
 * weight gradient straight from the GRU output, without materialising the shifted copy (shift = 0: agnn_split_f16_dropout). */
int agnn_split_f16_shifted(const float* x, int64_t rows, int64_t cols, int64_t ld_x, const float* amax, void* hi, void* lo,
                           int64_t ld_out, float dropout_p, const uint64_t* rng_state, uint32_t rng_stream,
                           int32_t seq_len, int32_t shift, agnn_stream_t stream);
/* The fused weight of one destination type of a PyG HeteroConv{SAGEConv} layer from the k relations' parameters
 * (contiguous fp32 [n, f] / [n]): wcat [n, (k + 1) f] = [sum_r lin_r_r | lin_l_1 | .. | lin_l_k] * scale,
 * bias [n] = sum_r bias_l_r * scale -- one launch instead of stack / sum / cat per layer and type. */
int agnn_sage_weights(int k, int32_t n, int32_t f, const float* const* lin_r /* host [k] */,
                      const float* const* lin_l, const float* const* bias_l, float scale, float* wcat, float* bias,
                      agnn_stream_t stream);
int agnn_l2norm_relu_fwd(const float* x, int64_t ld_x, float* y, int64_t ld_y, float* inv_norm, int64_t rows, int cols,
                         int relu_first, float eps, agnn_stream_t stream);
int agnn_l2norm_relu_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const float* inv_norm, float* dx,
                         int64_t ld_dx, int64_t rows, int cols, int relu_first, agnn_stream_t stream);
int agnn_colsum_partials(const float* x, int64_t ld_x, float* partials, float* out /* optional [cols] */, int64_t rows,
                         int cols, agnn_stream_t stream);
/* One pass over the gradient g [rows, cols] entering the backward of a projection (nn.Linear backward; the ReLU
 * in front of it when relu_out, the saved ReLU output, is given): g' = g * [relu_out > 0]; hi / lo = the TF32 operand
 * pair of g' (as agnn_split_tf32); partials[agnn_row_blocks(rows)][cols] and colsum[cols] = column sums of g' (the
 * bias gradient) when the two optional buffers are given. */
int agnn_grad_prepare(const float* g, int64_t ld_g, const float* relu_out /* optional */, int64_t ld_o, float* hi,
                      float* lo, int64_t ld_s, float* partials /* optional */, float* colsum /* optional */,
                      int64_t rows, int cols, agnn_stream_t stream);

/* agnn_grad_prepare with the operand pair in the F16X3 form: hi / lo are fp16 matrices (ld_s in fp16 elements), scaled
 * by the power of two derived from *amax (agnn_amax of g; the ReLU mask only removes entries). */
int agnn_grad_prepare_f16(const float* g, int64_t ld_g, const float* relu_out /* optional */, int64_t ld_o,
                          const float* amax, void* hi, void* lo, int64_t ld_s, float* partials /* optional */,
                          float* colsum /* optional */, int64_t rows, int cols, agnn_stream_t stream);

/* ------------------------------------------------------------ objective and lookup-table gradients
 * agnn_softmax_ce_*: nn.CrossEntropyLoss(ignore_index, label_smoothing) with mean reduction over the rows
 * that are not ignored -- one call per task head; MultiTaskLoss sums them (analysisgnn/models/analysis.py:
 * 881-908, 1035-1037; label_smoothing = 0.1 at :893).  Labels are int64 in [0, cols) or ignore_index.
 * fwd writes lse[rows] (log-sum-exp per row, kept for the backward), partials[agnn_ce_blocks(rows)][2] and
 * out[2] = {mean loss, number of rows that count}; bwd writes
 *   dlogits = grad_out / rows_that_count * (softmax - (1 - smoothing) onehot - smoothing / cols)   (0 for ignored rows).
 * agnn_embedding_bwd: d weight of a small nn.Embedding (pitch spelling 35 x 64, key signature 15 x 64,
 * analysisgnn/models/analysis.py:427-428, 572-574): dweight[e] = sum over rows with idx == e of g[row], summed
 * in a fixed order; n_emb * dim <= 3072; partials[agnn_embedding_bwd_blocks(rows)][n_emb * dim].
 */
int agnn_ce_blocks(int64_t rows);
int agnn_softmax_ce_fwd(const float* logits, int64_t ld, const int64_t* labels, int64_t rows, int cols, float smoothing,
                        int64_t ignore_index, float* lse, float* partials, float* out, agnn_stream_t stream);
int agnn_softmax_ce_bwd(const float* logits, int64_t ld, const int64_t* labels, const float* lse, int64_t rows, int cols,
                        float smoothing, int64_t ignore_index, const float* out, const float* grad_out /* device scalar */,
                        float* dlogits, int64_t ld_d, agnn_stream_t stream);
/* Same backward writing rows of `cols_padded` >= cols columns (the extra columns are zeroed: a 16-byte aligned
 * gradient for the head's backward GEMMs without a padding copy) and, optionally, *amax_out = max(*amax_out,
 * |grad_out / rows_that_count|), an upper bound of every |dlogits| entry (|softmax - target| <= 1). */
int agnn_softmax_ce_bwd_padded(const float* logits, int64_t ld, const int64_t* labels, const float* lse, int64_t rows,
                               int cols, float smoothing, int64_t ignore_index, const float* out, const float* grad_out,
                               float* dlogits, int64_t ld_d, int cols_padded, float* amax_out, agnn_stream_t stream);
int agnn_embedding_bwd_blocks(int64_t rows);
int agnn_embedding_bwd(const float* g, int64_t ld_g, const int64_t* idx, int64_t rows, int dim, int n_emb,
                       float* partials, float* dweight, agnn_stream_t stream);

/* ------------------------------------------------------------ onset-wise decode
 * Replaces onsetwise_logit_aggregation (analysisgnn/models/analysis.py:44-101, called from predict at :1588); the
 * onset mean itself (:66) is agnn_gather_reduce with the self term.  All buffers are device memory.
 *
 * agnn_softmax2_rows   y[i] = softmax(softmax(x[rows ? rows[i] : i])): the softmax after the onset mean (:66) and the
 *                      one after the valid-label selection (:68) in one pass; cols <= 256.
 * agnn_run_heads       heads[u] = first position of the u-th run of equal consecutive keys, *n_runs = number of runs
 *                      (torch.unique + the run starts of its inverse, :79-81, for keys that do not decrease;
 *                      *unsorted is set to 1 if they do).  With n_dev only the first min(n, *n_dev) keys count.
 * agnn_row_argmax      out[u] = first arg-max of row rowmap[heads[u]] (either map may be null), u < min(n_max, *n_dev).
 * agnn_decode_assign   the loop over change points (:85-99): with ov(i) = onsets_f[onset_heads[cp_heads[i]]], every row j
 *                      whose onsets[j] lies in [ov(i), ov(i+1)) for some i < *n_cp - 1 is overwritten by row
 *                      rowmap[onset_heads[cp_heads[i]]] of the same matrix (the last segment stays as it is).
 */
int agnn_softmax2_rows(const float* x, int64_t ld_x, const int32_t* rows /* optional */, int64_t n_out, int cols,
                       float* y, int64_t ld_y, agnn_stream_t stream);
size_t agnn_run_heads_workspace(int64_t n);
int agnn_run_heads(const int64_t* keys, int64_t n, const int32_t* n_dev /* optional */, int32_t* heads /* [n] */,
                   int32_t* n_runs /* device */, int32_t* unsorted /* optional, device, pre-zeroed */, void* workspace,
                   size_t workspace_bytes, agnn_stream_t stream);
int agnn_row_argmax(const float* x, int64_t ld, const int32_t* heads /* optional */, const int32_t* rowmap /* optional */,
                    const int32_t* n_dev /* optional */, int64_t n_max, int cols, int64_t* out, agnn_stream_t stream);
int agnn_decode_assign(float* y, int64_t ld, int cols, const int64_t* onsets, int64_t n_rows, const int64_t* onsets_f,
                       const int32_t* onset_heads, const int32_t* rowmap /* optional */, const int32_t* cp_heads,
                       const int32_t* n_cp /* device */, agnn_stream_t stream);

/* ------------------------------------------------------------ score-graph construction
 * Replaces hetero_graph_from_note_array (analysisgnn/utils/hgraph.py:214-300; rest_array=None,
 * pot_edge_dist=0) for a batch of scores: onset (0) / consecutive (1) / during (2) / rest (3) edges in
 * the reference's emission order, node ids offset by the score's first note.  Notes of a score must be
 * sorted by onset.  score_ptr [S+1]: first note of every score; key_base [S+1]: prefix sum of
 * (max_end - first_onset + 1) per score (host arithmetic on the note arrays), key_slots = key_base[S].
 * edges: int64 [3][capacity] (src, dst, type rows); *n_edges receives the number of edges the scores
 * have -- if it exceeds `capacity` the surplus was not written and the call must be repeated.
 */
size_t agnn_score_graph_workspace(int32_t n_notes, int32_t n_scores, int64_t key_slots);
int agnn_score_graph_build(int32_t n_scores, const int32_t* score_ptr, const int32_t* key_base, const int32_t* onset,
                           const int32_t* duration, int32_t n_notes, int64_t key_slots, int64_t* edges,
                           int64_t capacity, int32_t* n_edges, void* workspace, size_t workspace_bytes,
                           agnn_stream_t stream);

/* ------------------------------------------------------------ subgraph sampling
 * Replaces the CPU sampling / collation behind the reference's loaders (analysisgnn/data/datamodules/
 * analysis.py:270-323 -> graphmuse MuseNeighborLoader -> PyG NeighborSampler -> pyg-lib, third party) and the
 * in-tree window step (analysisgnn/data/datasets/chord.py:217-229, analysisgnn/utils/hgraph.py:404-452).
 * Semantics = oracle/graph.py (window_subgraph, neighbor_sample); bit-identical for a given seed.
 *
 * agnn_window_subgraph: for batch slot b, the node-induced subgraph of the contiguous window
 * [node_lo[b], node_lo[b] + win_size[b]) of one score: its candidate edges are the corpus edges
 * [edge_lo[b], edge_lo[b] + cand_ptr[b+1] - cand_ptr[b]); kept edges (both ends inside) come out in corpus
 * order, re-indexed to out_off[b] + (id - node_lo[b]) -- the collated batch.  edges: int64 [3][capacity].
 */
size_t agnn_window_workspace(int64_t n_cand);
int agnn_window_subgraph(int32_t n_slots, int64_t n_cand, const int64_t* cand_ptr, const int64_t* edge_lo,
                         const int64_t* node_lo, const int32_t* win_size, const int64_t* out_off, const int64_t* src,
                         const int64_t* dst, const int64_t* type /* optional */, int64_t* edges,
                         int64_t* edge_id /* optional [capacity] */, int64_t capacity, int32_t* n_out, void* workspace,
                         size_t workspace_bytes, agnn_stream_t stream);

/* k-hop uniform neighbour sampling on a relation-major CSR keyed on the destination (agnn_csr_build),
 * one node type.  local[n_nodes] / nodes[] hold the batch-local id of every node (or -1) and the
 * discovered nodes in discovery order.  Per hop: _count (edges per relation x frontier node, scanned),
 * the caller reads counts[n_rel * frontier_n] = n_cand, then _draw (sampling without replacement by
 * partial Fisher-Yates with draw t = rng(seed, hop, relation, destination, t); all neighbours when
 * degree <= fanout or fanout < 0; new sources are appended to `nodes` in the order the sequential
 * algorithm meets them).  Candidates are ordered (relation, frontier node, draw): cand_slot = CSR slot,
 * cand_src = global source, cand_dst = local destination, src_local = local source.  After _draw,
 * flag[n_cand] holds the number of newly discovered nodes.  fanout <= 32.
 */
int agnn_sample_init(int32_t n_nodes, const int32_t* seeds, int32_t n_seeds, int32_t* local, int32_t* nodes,
                     agnn_stream_t stream);
size_t agnn_sample_hop_workspace(int32_t n_rel, int32_t frontier_n);
int agnn_sample_hop_count(int32_t n_rel, int32_t n_nodes, const int32_t* rowptr, const int32_t* nodes,
                          int32_t frontier_lo, int32_t frontier_n, int32_t fanout, int32_t* counts, void* workspace,
                          size_t workspace_bytes, agnn_stream_t stream);
int agnn_sample_hop_draw(int32_t n_rel, int32_t n_nodes, const int32_t* rowptr, const int32_t* col, int32_t* nodes,
                         int32_t frontier_lo, int32_t frontier_n, int32_t n_known, int32_t hop, int32_t fanout,
                         uint64_t seed, const int32_t* counts, int32_t n_cand, int32_t* cand_slot, int32_t* cand_src,
                         int32_t* cand_dst, int32_t* src_local, int32_t* local, int32_t* first_pos, int32_t* flag,
                         void* workspace, size_t workspace_bytes, agnn_stream_t stream);

/* ------------------------------------------------------------ GRU recurrence
 * Replaces the sequential part of nn.GRU in the sequence branches (analysisgnn/models/cadence.py:249-260,
 * 276-285; analysisgnn/models/analysis.py:527-537; MetricalConvLayer.seq, analysisgnn/models/core/gnn.py:498,
 * 523).  The input projections gi = X W_ih^T + b_ih of all time steps and every weight / input gradient are
 * agnn_gemm calls made by the caller; these kernels run the time loop (one layer, n_dir = 1 or 2 directions, h0 = 0,
 * batch-first [B, T, .]): hidden size 32, 64 or 128 with W_hh resident in registers for all T steps
 * (AGNN_GRU_RESIDENT: one launch), multiples of 64 from 192 to 2048 -- MetricalConvLayer's GRU(512, 512) -- as one
 * launch per time step over (unit block, sequence block, direction) tiles (AGNN_GRU_STEPWISE; agnn_gru_fwd takes both,
 * the backward of the second kind is agnn_gru_bwd_stepwise).
 * Pointer-array arguments are HOST arrays of n_dir device pointers.
 *   fwd:  out [B, T, n_dir*H]; gates[d] [B, T, 4H] (r, z, n, W_hn h + b_hn) kept for the backward (may be NULL
 *         arrays for inference).
 *   bwd:  from dout, out, gates: dgi[d] = dL/d(gi_d) [B, T, 3H] and dgh[d] = dL/d(W_hh h + b_hh) [B, T, 3H];
 *         then db_ih = colsum(dgi), db_hh = colsum(dgh), dW_ih = dgi^T X, dW_hh = dgh^T H_prev, dX = sum_d dgi_d W_ih_d.
 */
#define AGNN_GRU_RESIDENT 1
#define AGNN_GRU_STEPWISE 2
int agnn_gru_supported(int hidden); /* 0, AGNN_GRU_RESIDENT or AGNN_GRU_STEPWISE */
/* Hidden size 128 has two resident implementations: the SIMT kernels (2 sequences per CTA, W_hh in registers as fp32)
 * and tensor-core kernels (8 sequences per CTA, W_hh as fp16 hi / lo MMA fragments, 3 MMAs per product).  mode 0: SIMT
 * for both passes, 1: tensor cores for both, 2: forward on the tensor cores when the SIMT kernel would hold more than half
 * of the SMs, i.e. more than 148 (sequence, direction) pairs (default; initial value from the environment variable
 * AGNN_GRU_TC), 3: backward only.  Returns the previous mode; any other argument only queries. */
int agnn_gru_mode(int mode);
int agnn_gru_fwd(int32_t batch, int32_t steps, int32_t hidden, int32_t n_dir, const float* const* gi /* host */,
                 const float* const* w_hh /* host */, const float* const* b_hh /* host */, float* out,
                 float* const* gates /* host, optional */, agnn_stream_t stream);
int agnn_gru_bwd(int32_t batch, int32_t steps, int32_t hidden, int32_t n_dir, const float* const* w_hh /* host */,
                 const float* out, const float* const* gates /* host */, const float* dout,
                 float* const* dgi /* host */, float* const* dgh /* host */, agnn_stream_t stream);
/* Same backward; amax[d] (optional host array of n_dir device scalars, zeroed by the caller) receives max |dgi[d]|,
 * which also bounds dgh[d] (the n gate of dgh is the one of dgi times r, |r| <= 1): the fp16 operand scale of both for
 * the weight-gradient products, without a pass over them. */
int agnn_gru_bwd_amax(int32_t batch, int32_t steps, int32_t hidden, int32_t n_dir, const float* const* w_hh,
                      const float* out, const float* const* gates, const float* dout, float* const* dgi,
                      float* const* dgh, float* const* amax, agnn_stream_t stream);
/* Backward of an AGNN_GRU_STEPWISE hidden size: same outputs as agnn_gru_bwd_amax.  w_hh_t[d] = W_hh^T of direction d,
 * [H, 3H] row-major (the carried gradient dgh W_hh reads it K-major); carry = n_dir * B * H floats of scratch (z * dh
 * handed from one time step's launch to the next). */
int agnn_gru_bwd_stepwise(int32_t batch, int32_t steps, int32_t hidden, int32_t n_dir,
                          const float* const* w_hh_t /* host */, const float* out, const float* const* gates /* host */,
                          const float* dout, float* const* dgi /* host */, float* const* dgh /* host */,
                          float* const* amax /* host, optional */, float* carry, agnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AGNN_H_ */
