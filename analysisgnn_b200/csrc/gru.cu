// Recurrent part of the bidirectional GRU sequence branches.
//
// The reference runs nn.GRU over every score's padded note (or beat / measure) sequence: the hybrid
// branch of the encoder (analysisgnn/models/cadence.py:249-260, 276-285; analysisgnn/models/analysis.py:
// 527-537) and MetricalConvLayer.seq (analysisgnn/models/core/gnn.py:498, 523).  cuDNN's fp32 GRU (TF32
// is off for parity) became the critical path of the step: 11 of 17.5 ms.  Here the GRU is split the
// way its arithmetic allows:
//   * everything that is a big GEMM -- the input projections X W_ih^T of ALL time steps and, in the
//     backward, dX, dW_ih, dW_hh -- runs on agnn_gemm (tensor cores, 3xTF32);
//   * only the truly sequential part stays in these two kernels: per time step the H x 3H recurrent
//     mat-vec, the gate nonlinearities and (backward) the carried dh.
// Sequences are independent, so there is no inter-CTA synchronisation at all: a CTA owns kSeq sequences
// of one direction for their whole length, 3H threads, and every thread keeps its 4 x H/4 block of W_hh
// in REGISTERS for all T steps -- the weights are read from memory once.  The per-step inputs are copied
// kPrefetch steps ahead into a shared-memory ring with cp.async (no registers, no stall on the dependent
// chain).  PyTorch gate order (r, z, n):
//   r = s(gi_r + gh_r)  z = s(gi_z + gh_z)  n = tanh(gi_n + r * gh_n)  h' = (1 - z) n + z h,  gh = W_hh h + b_hh
#include <cstring>

#include "common.cuh"

namespace agnn {
namespace {

#ifndef AGNN_GRU_SEQ
#define AGNN_GRU_SEQ 2
#endif
constexpr int kSeq = AGNN_GRU_SEQ;   // sequences per CTA (reuse of the register-resident weights; <= 3: 3H threads own kSeq * H units)
constexpr int kPrefetch = 4;  // time steps of input prefetched ahead

struct GruParams {
  int batch, steps, n_dir;
  // per direction d: gi [B, T, 3H] (input projections + b_ih), w_hh [3H, H], b_hh [3H]
  const float* gi[2];
  const float* w_hh[2];
  const float* b_hh[2];
  float* out;          // [B, T, n_dir * H]
  int64_t ld_out;      // n_dir * H
  float* gates[2];     // [B, T, 4H]: r, z, n, gh_n  (saved for the backward)
  // backward
  const float* dout;   // [B, T, n_dir * H]
  float* dgi[2];       // [B, T, 3H]
  float* dgh[2];       // [B, T, 3H]
  float* amax[2];      // optional per direction: max |dgi| (>= max |dgh|: dgh = dgi with the n gate times r, |r| <= 1)
  // wide hidden sizes (one launch per time step, see below)
  int hidden;
  const float* w_hh_t[2];  // [H, 3H] = W_hh^T (backward)
  float* keep;             // [n_dir, B, H]: z_t * dh_t carried to the next backward step
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Both kernels block a [H x H] piece of W_hh as 4 x (H/4) per thread: a thread owns 4 output
// elements and a quarter of the reduction range, so it needs only H/4 values of the broadcast vector per
// sequence and step (H/16 LDS.128 instead of H/4 -- the loop was shared-memory-latency bound with whole
// rows per thread), keeps 8 independent accumulation chains, and the 4 partial sums of an output meet in
// 2 xor-shuffles.  The broadcast vector lives in shared memory as 4 slices padded by 4 floats, which puts
// the 4 addresses a warp reads at once into different banks.
constexpr int kSlices = 4;

__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
               "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int H>
__global__ void __launch_bounds__(3 * H, 1) gru_fwd_kernel(const __grid_constant__ GruParams p) {
  constexpr int KS = H / kSlices;                // reduction range of a thread
  constexpr int LD = KS + 4;                     // padded slice
  __shared__ __align__(16) float h_s[kSeq][kSlices][LD];
  __shared__ float gh_s[kSeq][3 * H];
  __shared__ float gi_s[kPrefetch][kSeq][3 * H];  // input projections of the next steps (cp.async ring)
  const int j = threadIdx.x;
  const int c = j & 3, g = j >> 2;               // slice, row group: rows 4g .. 4g+3 of W_hh
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * kSeq;
  const int T = p.steps;
  float w[4][KS];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float* wr = p.w_hh[dir] + (int64_t)(4 * g + r) * H + c * KS;
#pragma unroll
    for (int k = 0; k < KS; k += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(wr + k));
      w[r][k] = t.x; w[r][k + 1] = t.y; w[r][k + 2] = t.z; w[r][k + 3] = t.w;
    }
  }
  const float bias = __ldg(p.b_hh[dir] + 4 * g + c);   // of the row this lane publishes
  for (int i = j; i < kSeq * kSlices * LD; i += 3 * H) (&h_s[0][0][0])[i] = 0.f;
  auto time_of = [&](int step) { return dir == 0 ? step : T - 1 - step; };
  auto fetch_gi = [&](int step, int slot) {      // every thread copies its column of each sequence's row
    if (step < T) {
      const int t = time_of(step);
#pragma unroll
      for (int s = 0; s < kSeq; ++s) {
        const int b = b0 + s;
        if (b < p.batch) cp_async4(&gi_s[slot][s][j], p.gi[dir] + ((int64_t)b * T + t) * (3 * H) + j);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int q = 0; q < kPrefetch; ++q) fetch_gi(q, q);
  // gate phase: thread j < kSeq * H owns hidden unit u of sequence sq
  const int sq = j / H, u = j - sq * H;
  const bool owner = j < kSeq * H && b0 + sq < p.batch;
  float* h_own = &h_s[owner ? sq : 0][u / KS][u % KS];
  __syncthreads();
  for (int step0 = 0; step0 < T; step0 += kPrefetch) {
#pragma unroll
    for (int q = 0; q < kPrefetch; ++q) {
      const int step = step0 + q;
      if (step >= T) break;
      // 1. recurrent mat-vec, partial over this thread's slice
      float acc[4][kSeq];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < kSeq; ++s) acc[r][s] = 0.f;
#pragma unroll
      for (int k = 0; k < KS; k += 4) {
#pragma unroll
        for (int s = 0; s < kSeq; ++s) {
          const float4 hv = *reinterpret_cast<const float4*>(&h_s[s][c][k]);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            acc[r][s] = fmaf(w[r][k], hv.x, acc[r][s]);
            acc[r][s] = fmaf(w[r][k + 1], hv.y, acc[r][s]);
            acc[r][s] = fmaf(w[r][k + 2], hv.z, acc[r][s]);
            acc[r][s] = fmaf(w[r][k + 3], hv.w, acc[r][s]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < kSeq; ++s) {
          acc[r][s] += __shfl_xor_sync(0xffffffffu, acc[r][s], 1);
          acc[r][s] += __shfl_xor_sync(0xffffffffu, acc[r][s], 2);
        }
#pragma unroll
      for (int s = 0; s < kSeq; ++s) {
        const float mine = c == 0 ? acc[0][s] : c == 1 ? acc[1][s] : c == 2 ? acc[2][s] : acc[3][s];
        gh_s[s][4 * g + c] = mine + bias;
      }
      cp_async_wait<kPrefetch - 1>();            // this step's gi has landed (own copies; the barrier publishes)
      __syncthreads();
      // 2. gates and the new hidden state
      if (owner) {
        const int t = time_of(step);
        const float ghn = gh_s[sq][2 * H + u];
        const float r = sigmoidf_(gi_s[q][sq][u] + gh_s[sq][u]);
        const float z = sigmoidf_(gi_s[q][sq][H + u] + gh_s[sq][H + u]);
        const float n = tanhf(gi_s[q][sq][2 * H + u] + r * ghn);
        const float hn = (1.f - z) * n + z * *h_own;
        *h_own = hn;
        const int64_t row = (int64_t)(b0 + sq) * T + t;
        p.out[row * p.ld_out + dir * H + u] = hn;
        if (p.gates[dir]) {
          float* gt = p.gates[dir] + row * (4 * H);
          gt[u] = r; gt[H + u] = z; gt[2 * H + u] = n; gt[3 * H + u] = ghn;
        }
      }
      __syncthreads();
      fetch_gi(step + kPrefetch, q);             // slot q was consumed before the barrier above
    }
  }
}

template <int H>
__global__ void __launch_bounds__(3 * H, 1) gru_bwd_kernel(const __grid_constant__ GruParams p) {
  constexpr int KS = H / kSlices;
  constexpr int LD = KS + 4;
  __shared__ __align__(16) float dgh_s[kSeq][3][kSlices][LD];
  __shared__ float part_s[kSeq][3][H];
  __shared__ float in_s[kPrefetch][kSeq][6][H];  // dout, r, z, n, gh_n, h_prev of the next steps
  const int j = threadIdx.x;
  float gmax = 0.f;
  const int gate = j / H, m = j - gate * H;
  const int rs = m & 3, cg = m >> 2;             // row slice, column group: columns 4cg .. 4cg+3 of the gate block
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * kSeq;
  const int T = p.steps;
  float w[KS][4];                                // w[i][cc] = W_hh[gate*H + rs*KS + i, 4cg + cc]
#pragma unroll
  for (int i = 0; i < KS; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p.w_hh[dir] + (int64_t)(gate * H + rs * KS + i) * H + 4 * cg));
    w[i][0] = t.x; w[i][1] = t.y; w[i][2] = t.z; w[i][3] = t.w;
  }
  // processing order is the reverse of the forward's: forward-direction t = T-1 .. 0, reverse t = 0 .. T-1
  auto time_of = [&](int step) { return dir == 0 ? T - 1 - step : step; };
  auto fetch_in = [&](int step, int slot) {
    if (step < T) {
      const int t = time_of(step);
      const int tp = dir == 0 ? t - 1 : t + 1;   // time of h_{prev} in the forward recurrence
      for (int e = j; e < kSeq * 6 * H; e += 3 * H) {
        const int s = e / (6 * H), f = (e - s * 6 * H) / H, uu = e % H;
        const int b = b0 + s;
        if (b >= p.batch) continue;
        const int64_t row = (int64_t)b * T + t;
        const float* src;
        if (f == 0) src = p.dout + row * p.ld_out + dir * H + uu;
        else if (f < 5) src = p.gates[dir] + row * (4 * H) + (f - 1) * H + uu;
        else {
          if (tp < 0 || tp >= T) continue;       // h_prev = 0: handled at the consumer
          src = p.out + ((int64_t)b * T + tp) * p.ld_out + dir * H + uu;
        }
        cp_async4(&in_s[slot][s][f][uu], src);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int q = 0; q < kPrefetch; ++q) fetch_in(q, q);
  const int sq = j / H, u = j - sq * H;          // gate-phase ownership: unit u of sequence sq
  const bool owner = j < kSeq * H && b0 + sq < p.batch;
  float carry = 0.f;                             // dL/dh carried to the previous step
  for (int i = j; i < kSeq * 3 * kSlices * LD; i += 3 * H) (&dgh_s[0][0][0][0])[i] = 0.f;
  cp_async_wait<kPrefetch - 1>();
  __syncthreads();
  for (int step0 = 0; step0 < T; step0 += kPrefetch) {
#pragma unroll
    for (int q = 0; q < kPrefetch; ++q) {
      const int step = step0 + q;
      if (step >= T) break;
      float keep = 0.f;                          // dh_total * z, the direct path to h_prev
      if (owner) {
        const int t = time_of(step);
        const int tp = dir == 0 ? t - 1 : t + 1;
        const float gd = in_s[q][sq][0][u], r = in_s[q][sq][1][u], z = in_s[q][sq][2][u], n = in_s[q][sq][3][u];
        const float ghn = in_s[q][sq][4][u];
        const float hp = (tp >= 0 && tp < T) ? in_s[q][sq][5][u] : 0.f;
        const float dh = gd + carry;
        const float dn_pre = dh * (1.f - z) * (1.f - n * n);
        const float dz_pre = dh * (hp - n) * z * (1.f - z);
        const float dr_pre = dn_pre * ghn * r * (1.f - r);
        keep = dh * z;
        gmax = fmaxf(gmax, fmaxf(fabsf(dr_pre), fmaxf(fabsf(dz_pre), fabsf(dn_pre))));
        const int64_t row = ((int64_t)(b0 + sq) * T + t) * (3 * H);
        p.dgi[dir][row + u] = dr_pre; p.dgi[dir][row + H + u] = dz_pre; p.dgi[dir][row + 2 * H + u] = dn_pre;
        const float dghn = dn_pre * r;
        p.dgh[dir][row + u] = dr_pre; p.dgh[dir][row + H + u] = dz_pre; p.dgh[dir][row + 2 * H + u] = dghn;
        dgh_s[sq][0][u / KS][u % KS] = dr_pre;
        dgh_s[sq][1][u / KS][u % KS] = dz_pre;
        dgh_s[sq][2][u / KS][u % KS] = dghn;
      }
      __syncthreads();                           // (A)
      // dh_prev[k] += sum_i dgh[gate*H + i] * W_hh[gate*H + i, k]: partial over this thread's row slice
      float acc[4][kSeq];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int s = 0; s < kSeq; ++s) acc[cc][s] = 0.f;
#pragma unroll
      for (int i = 0; i < KS; i += 4) {
#pragma unroll
        for (int s = 0; s < kSeq; ++s) {
          const float4 dv = *reinterpret_cast<const float4*>(&dgh_s[s][gate][rs][i]);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            acc[cc][s] = fmaf(w[i][cc], dv.x, acc[cc][s]);
            acc[cc][s] = fmaf(w[i + 1][cc], dv.y, acc[cc][s]);
            acc[cc][s] = fmaf(w[i + 2][cc], dv.z, acc[cc][s]);
            acc[cc][s] = fmaf(w[i + 3][cc], dv.w, acc[cc][s]);
          }
        }
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int s = 0; s < kSeq; ++s) {
          acc[cc][s] += __shfl_xor_sync(0xffffffffu, acc[cc][s], 1);
          acc[cc][s] += __shfl_xor_sync(0xffffffffu, acc[cc][s], 2);
        }
#pragma unroll
      for (int s = 0; s < kSeq; ++s)
        part_s[s][gate][4 * cg + rs] = rs == 0 ? acc[0][s] : rs == 1 ? acc[1][s] : rs == 2 ? acc[2][s] : acc[3][s];
      fetch_in(step + kPrefetch, q);             // slot q was consumed before barrier (A)
      cp_async_wait<kPrefetch - 1>();            // the next step's inputs have landed; barrier (B) publishes them
      __syncthreads();                           // (B)
      if (owner) carry = keep + (part_s[sq][0][u] + part_s[sq][1][u]) + part_s[sq][2][u];
      // no third barrier: dgh_s is rewritten only after barrier (B), which every thread passes after its
      // mat-vec reads; part_s is rewritten only after the next barrier (A), which follows these reads
    }
  }
  if (p.amax[blockIdx.y]) {                      // operand scale of dgi / dgh for the weight-gradient GEMMs
    uint32_t m = __float_as_uint(gmax);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(reinterpret_cast<unsigned int*>(p.amax[blockIdx.y]), m);
  }
}


// ---------------------------------------------------------------------------------------------------------
// Wide hidden sizes (multiples of 64 above 128: MetricalConvLayer.seq runs nn.GRU(512, 512) in config 4,
// analysisgnn/models/core/gnn.py:498).  W_hh no longer fits the registers of one SM (3 MB at H = 512) and 64
// sequences cannot occupy 148 SMs by sequence alone, so a time step becomes ONE launch over (unit block, sequence
// block, direction) tiles and the step order is the launch order: no grid barrier, no spinning, and the whole time
// loop is part of the captured step graph.  Every CTA computes a small product
//     C[TS sequences, TN columns] = A[TS, K] . B[TN, K]^T
// with both operands K-major fp32, streamed from L2 through a cp.async ring in 64-float chunks:
//   forward   A = h_{t-1} (rows of `out`),               B = the r, z, n rows of 8 hidden units of W_hh, K = H;
//             epilogue = the gates and new state of those 8 units;
//   backward  A = dgh of the step processed before,      B = 16 rows of W_hh^T,                        K = 3H;
//             epilogue = dh, the gate derivatives of those 16 units and the carried z * dh.
// The product runs on the tensor cores at fp32 accuracy: mma.sync m16n8k8 TF32 with every operand split in registers
// into hi = rna(x), lo = rna(x - hi) and D += A_lo B_hi + A_hi B_lo + A_hi B_hi (the dropped lo x lo term is 2^-22
// relative).  The tensor core adds into its fp32 accumulator with truncation, so a chain ends after every chunk
// (12 / 6 MMAs) and the chains are summed with ordinary fp32 adds -- the same rule as the TMEM chains of gemm.cu.
// (A first version did this product with packed FFMA2 on 4 x 6 register tiles: shared-memory bound at 18 us per
// step -- measured: LDS.128 costs 4.2 cycles per warp whatever the broadcast pattern, FFMA2 issues at half the FFMA
// rate -- against 60 us per step for the library RNN; fragments read with conflict-free LDS.32 need a sixth of the
// shared-memory cycles.)  A warp owns one 8-wide k step of every chunk and the whole tile; the 8 partial tiles meet
// in shared memory, then thread (sequence, unit pair) finishes the epilogue.
// Bytes a step pulls from L2 (H = 512, 64 sequences, 2 directions): forward 128 CTAs x (128 KB of h + 48 KB of W_hh),
// backward 128 x (192 KB of dgh + 96 KB of W_hh^T).
namespace wide {
constexpr int kThreads = 256;
constexpr int KC = 64;            // floats of K per pipeline stage
constexpr int kStages = 4;
constexpr int LD = KC + 4;        // row stride = 4 banks: the 8 rows x 4 columns a fragment load touches cover all 32

// Programmatic dependent launch (launch attribute set by launch_steps): the next time step's grid may be scheduled as
// soon as every CTA of this one has passed launch_dependents(), and runs its independent prologue until
// grid_dependency_wait(), which returns when the previous grid has completed and its writes are visible.  Without
// the attribute both are no-ops.  Hides the launch latency and the cold epilogue loads of a 10-us kernel.
__device__ __forceinline__ void launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(x - __uint_as_float(hi)));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct Operands {
  const float* a;       // tile row 0 of A
  int64_t a_stride;     // floats between A rows
  int a_rows;           // valid rows (the rest read as zero)
  const float* b;       // B tile row l lives at b + (l / BG) * b_group_stride + (l % BG) * b_row_stride
  int64_t b_group_stride, b_row_stride;
  int k;                // multiple of KC
};

constexpr int KQ = kThreads / 32;  // every warp owns one 8-wide k step of a chunk (KC / 8 == KQ) and the WHOLE tile

template <int TS, int TN>
constexpr size_t smem_bytes() {
  constexpr size_t ring = (size_t)kStages * (TS + TN) * LD, part = (size_t)KQ * TS * (TN + 1);
  return (ring > part ? ring : part) * sizeof(float);
}

// c_s[q][row][col] (row stride TN + 1, aliasing the ring) = the partial product of warp q's k steps; the caller sums
// over q after the barrier this function ends with.  Warp = k slice rather than row slab: an operand element is then
// split into hi / lo by exactly one warp (with row slabs the B fragments were split once per slab: the loop was
// issue bound, 5.4 instructions per MMA; now 3.4).
template <int TS, int TN, int BG>
__device__ __forceinline__ void tile_mma(float* smem, const Operands& op) {
  constexpr int MT = TS / 16, NT = TN / 8;
  static_assert(KC / 8 == KQ, "one k step per warp and chunk");
  constexpr int NA = TS * (KC / 4) / kThreads, NB = (TN * (KC / 4) + kThreads - 1) / kThreads;
  static_assert(TS * (KC / 4) % kThreads == 0, "A copies per thread");
  float* a_s = smem;
  float* b_s = smem + kStages * TS * LD;
  const int tid = threadIdx.x, lane = tid & 31, kq = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // this thread's copies of a chunk: fixed (row, 16-byte piece) pairs, the source advances by KC per chunk
  const float* a_src[NA];
  const float* b_src[NB];
  uint32_t a_dst[NA], b_dst[NB];
  bool a_ok[NA], b_ok[NB];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const int e = tid + i * kThreads, r = e / (KC / 4), f = e % (KC / 4);
    a_ok[i] = r < op.a_rows;
    a_src[i] = op.a + (a_ok[i] ? r * op.a_stride + 4 * f : 0);
    a_dst[i] = (uint32_t)__cvta_generic_to_shared(a_s + r * LD + 4 * f);
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int e = tid + i * kThreads, l = e / (KC / 4), f = e % (KC / 4);
    b_ok[i] = l < TN;
    b_src[i] = op.b + (b_ok[i] ? (l / BG) * op.b_group_stride + (l % BG) * op.b_row_stride + 4 * f : 0);
    b_dst[i] = (uint32_t)__cvta_generic_to_shared(b_s + l * LD + 4 * f);
  }
  auto issue = [&](int chunk, int slot) {
    const int k0 = chunk * KC;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int n = a_ok[i] ? 16 : 0;            // 0 source bytes: the 16 destination bytes are zero-filled
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a_dst[i] + slot * (TS * LD * 4)),
                   "l"(a_src[i] + (a_ok[i] ? k0 : 0)), "r"(n)
                   : "memory");
    }
#pragma unroll
    for (int i = 0; i < NB; ++i)
      if (b_ok[i])
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(b_dst[i] + slot * (TN * LD * 4)),
                     "l"(b_src[i] + k0)
                     : "memory");
  };
  float tot[MT][NT][4], acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) tot[mt][nt][i] = acc[mt][nt][i] = 0.f;
  const int n_chunks = op.k / KC;
#pragma unroll
  for (int q = 0; q < kStages - 1; ++q) {
    if (q < n_chunks) issue(q, q);
    cp_async_commit();
  }
  for (int kc = 0; kc < n_chunks; ++kc) {
    cp_async_wait<kStages - 2>();     // chunk kc has landed (own copies); the barrier publishes everyone's and
    __syncthreads();                  // retires the slot refilled below (read in iteration kc - 1)
    const int nxt = kc + kStages - 1;
    if (nxt < n_chunks) issue(nxt, nxt % kStages);
    cp_async_commit();
    const int slot = kc % kStages;
    const float* a_t = a_s + (slot * TS + g) * LD + kq * 8 + t;
    const float* b_t = b_s + (slot * TN + g) * LD + kq * 8 + t;
    uint32_t bh[NT][2], bl[NT][2];    // B fragments: (k = t, n = g) (k = t + 4, n = g)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      split_tf32(b_t[nt * 8 * LD], bh[nt][0], bl[nt][0]);
      split_tf32(b_t[nt * 8 * LD + 4], bh[nt][1], bl[nt][1]);
    }
    uint32_t ah[MT][4], al[MT][4];    // A fragments: (g, t) (g + 8, t) (g, t + 4) (g + 8, t + 4)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      split_tf32(a_t[mt * 16 * LD], ah[mt][0], al[mt][0]);
      split_tf32(a_t[(mt * 16 + 8) * LD], ah[mt][1], al[mt][1]);
      split_tf32(a_t[mt * 16 * LD + 4], ah[mt][2], al[mt][2]);
      split_tf32(a_t[(mt * 16 + 8) * LD + 4], ah[mt][3], al[mt][3]);
    }
    // term-major order: consecutive MMAs go to different accumulators (MT * NT apart on the same one) -- issued
    // accumulator by accumulator, each MMA waited for the one before it
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], al[mt], bh[nt][0], bh[nt][1]);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], ah[mt], bl[nt][0], bl[nt][1]);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], ah[mt], bh[nt][0], bh[nt][1]);
    if ((kc & 3) == 3 || kc == n_chunks - 1) {   // end the accumulation chain (12 MMAs): fp32 adds round to nearest
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            tot[mt][nt][i] += acc[mt][nt][i];
            acc[mt][nt][i] = 0.f;
          }
    }
  }
  __syncthreads();                    // every warp is done with the ring: the partial tiles take its place
  // C fragment: (g, 2t) (g, 2t + 1) (g + 8, 2t) (g + 8, 2t + 1)
  float* c_w = smem + ((size_t)kq * TS + g) * (TN + 1) + 2 * t;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      c_w[(mt * 16) * (TN + 1) + nt * 8] = tot[mt][nt][0];
      c_w[(mt * 16) * (TN + 1) + nt * 8 + 1] = tot[mt][nt][1];
      c_w[(mt * 16 + 8) * (TN + 1) + nt * 8] = tot[mt][nt][2];
      c_w[(mt * 16 + 8) * (TN + 1) + nt * 8 + 1] = tot[mt][nt][3];
    }
  __syncthreads();
}

constexpr int FTS = 64, FTN = 24;   // forward tile: 64 sequences x (3 gates x 8 units)
constexpr int BTS = 32, BTN = 16;   // backward tile: 32 sequences x 16 units

__global__ void __launch_bounds__(kThreads) gru_wide_fwd_kernel(const __grid_constant__ GruParams p, int step) {
  extern __shared__ __align__(16) float smem[];
  const float* c_s = smem;
  const int H = p.hidden, T = p.steps;
  const int dir = blockIdx.z, u0 = blockIdx.x * 8, row0 = blockIdx.y * FTS;
  const int t = dir == 0 ? step : T - 1 - step;
  const int tp = dir == 0 ? t - 1 : t + 1;       // time of h_{prev}
  const int bl = threadIdx.x >> 2, nq = threadIdx.x & 3;     // epilogue: sequence bl of the tile, units 2 nq, 2 nq + 1
  const int b = row0 + bl;
  launch_dependents();
  const int64_t row = (int64_t)b * T + t;
  // what the epilogue needs from global memory, requested before the product
  float gi[3][2], bias[3][2], hp[2];
  if (b < p.batch) {
#pragma unroll
    for (int gt = 0; gt < 3; ++gt) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(p.gi[dir] + row * (3 * H) + gt * H + u0 + 2 * nq));
      const float2 w = __ldg(reinterpret_cast<const float2*>(p.b_hh[dir] + gt * H + u0 + 2 * nq));
      gi[gt][0] = v.x; gi[gt][1] = v.y; bias[gt][0] = w.x; bias[gt][1] = w.y;
    }
  }
  // programmatic dependent launch: this grid was allowed to start while the previous time step was still running;
  // everything above is independent of it, everything below reads what it wrote
  grid_dependency_wait();
  hp[0] = hp[1] = 0.f;
  if (b < p.batch && step > 0) {
    const float2 v = *reinterpret_cast<const float2*>(p.out + ((int64_t)b * T + tp) * p.ld_out + dir * H + u0 + 2 * nq);
    hp[0] = v.x; hp[1] = v.y;
  }
  if (step > 0) {                                // h_0 = 0: the first step's product vanishes
    Operands op;
    op.a = p.out + ((int64_t)row0 * T + tp) * p.ld_out + dir * H;
    op.a_stride = (int64_t)T * p.ld_out;
    op.a_rows = p.batch - row0;
    op.b = p.w_hh[dir] + (int64_t)u0 * H;        // tile row l = gate * 8 + unit
    op.b_group_stride = (int64_t)H * H;          // gate block of W_hh
    op.b_row_stride = H;
    op.k = H;
    tile_mma<FTS, FTN, 8>(smem, op);
  }
  if (b >= p.batch) return;
  float gh[3][2];
#pragma unroll
  for (int gt = 0; gt < 3; ++gt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v = 0.f;
      if (step > 0) {
#pragma unroll
        for (int q = 0; q < KQ; ++q) v += c_s[((size_t)q * FTS + bl) * (FTN + 1) + gt * 8 + 2 * nq + j];
      }
      gh[gt][j] = v + bias[gt][j];
    }
  float hn[2], rr[2], zz[2], nn[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    rr[j] = sigmoidf_(gi[0][j] + gh[0][j]);
    zz[j] = sigmoidf_(gi[1][j] + gh[1][j]);
    nn[j] = tanhf(gi[2][j] + rr[j] * gh[2][j]);
    hn[j] = (1.f - zz[j]) * nn[j] + zz[j] * hp[j];
  }
  const int u = u0 + 2 * nq;
  *reinterpret_cast<float2*>(p.out + row * p.ld_out + dir * H + u) = make_float2(hn[0], hn[1]);
  if (p.gates[dir]) {
    float* gt = p.gates[dir] + row * (4 * H) + u;
    *reinterpret_cast<float2*>(gt) = make_float2(rr[0], rr[1]);
    *reinterpret_cast<float2*>(gt + H) = make_float2(zz[0], zz[1]);
    *reinterpret_cast<float2*>(gt + 2 * H) = make_float2(nn[0], nn[1]);
    *reinterpret_cast<float2*>(gt + 3 * H) = make_float2(gh[2][0], gh[2][1]);
  }
}

__global__ void __launch_bounds__(kThreads) gru_wide_bwd_kernel(const __grid_constant__ GruParams p, int step) {
  extern __shared__ __align__(16) float smem[];
  const float* c_s = smem;
  const int H = p.hidden, T = p.steps;
  const int dir = blockIdx.z, u0 = blockIdx.x * BTN, row0 = blockIdx.y * BTS;
  const int t = dir == 0 ? T - 1 - step : step;  // reverse of the forward's processing order
  const int tn = dir == 0 ? t + 1 : t - 1;       // the time the previous backward step processed
  const int tp = dir == 0 ? t - 1 : t + 1;       // time of h_{prev} in the forward recurrence
  const int bl = threadIdx.x >> 3, nq = threadIdx.x & 7;     // epilogue: sequence bl of the tile, units 2 nq, 2 nq + 1
  const int b = row0 + bl, u = u0 + 2 * nq;
  launch_dependents();
  const int64_t row = (int64_t)b * T + t;
  float gd[2], r[2], z[2], n[2], ghn[2], hp[2], kp[2];
  float* keep = p.keep + ((int64_t)dir * p.batch + b) * H + u;
  if (b < p.batch) {
    auto ld2 = [](const float* q, float (&o)[2]) { const float2 v = *reinterpret_cast<const float2*>(q); o[0] = v.x; o[1] = v.y; };
    const float* gt = p.gates[dir] + row * (4 * H) + u;
    ld2(p.dout + row * p.ld_out + dir * H + u, gd);
    ld2(gt, r); ld2(gt + H, z); ld2(gt + 2 * H, n); ld2(gt + 3 * H, ghn);
    hp[0] = hp[1] = kp[0] = kp[1] = 0.f;
    if (tp >= 0 && tp < T) ld2(p.out + ((int64_t)b * T + tp) * p.ld_out + dir * H + u, hp);
  }
  grid_dependency_wait();                        // the loads above do not depend on the previous backward step
  if (b < p.batch && step > 0) {
    const float2 v = *reinterpret_cast<const float2*>(keep);
    kp[0] = v.x; kp[1] = v.y;
  }
  if (step > 0) {                                // dh_prev[b, u] = sum_i dgh[b, i] W_hh[i, u]
    Operands op;
    op.a = p.dgh[dir] + ((int64_t)row0 * T + tn) * (3 * H);
    op.a_stride = (int64_t)T * 3 * H;
    op.a_rows = p.batch - row0;
    op.b = p.w_hh_t[dir] + (int64_t)u0 * 3 * H;
    op.b_group_stride = 0;
    op.b_row_stride = 3 * H;
    op.k = 3 * H;
    tile_mma<BTS, BTN, BTN>(smem, op);
  }
  float gmax = 0.f;
  if (b < p.batch) {
    float dr[2], dz[2], dn[2], dgn[2], kn[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v = 0.f;
      if (step > 0) {
#pragma unroll
        for (int q = 0; q < KQ; ++q) v += c_s[((size_t)q * BTS + bl) * (BTN + 1) + 2 * nq + j];
      }
      const float dh = step > 0 ? gd[j] + (kp[j] + v) : gd[j];
      dn[j] = dh * (1.f - z[j]) * (1.f - n[j] * n[j]);
      dz[j] = dh * (hp[j] - n[j]) * z[j] * (1.f - z[j]);
      dr[j] = dn[j] * ghn[j] * r[j] * (1.f - r[j]);
      dgn[j] = dn[j] * r[j];
      kn[j] = dh * z[j];
      gmax = fmaxf(gmax, fmaxf(fabsf(dr[j]), fmaxf(fabsf(dz[j]), fabsf(dn[j]))));
    }
    *reinterpret_cast<float2*>(keep) = make_float2(kn[0], kn[1]);
    float* gi_row = p.dgi[dir] + row * (3 * H) + u;
    float* gh_row = p.dgh[dir] + row * (3 * H) + u;
    *reinterpret_cast<float2*>(gi_row) = make_float2(dr[0], dr[1]);
    *reinterpret_cast<float2*>(gi_row + H) = make_float2(dz[0], dz[1]);
    *reinterpret_cast<float2*>(gi_row + 2 * H) = make_float2(dn[0], dn[1]);
    *reinterpret_cast<float2*>(gh_row) = make_float2(dr[0], dr[1]);
    *reinterpret_cast<float2*>(gh_row + H) = make_float2(dz[0], dz[1]);
    *reinterpret_cast<float2*>(gh_row + 2 * H) = make_float2(dgn[0], dgn[1]);
  }
  if (p.amax[dir]) {
    uint32_t m = __float_as_uint(gmax);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(reinterpret_cast<unsigned int*>(p.amax[dir]), m);
  }
}

inline bool supported(int hidden) { return hidden > 128 && hidden <= 2048 && hidden % 64 == 0; }

// One launch per time step, each a programmatic dependent of the one before: its launch latency and cold epilogue
// loads overlap the previous step's tail (forward 1.86 -> 1.63 ms, backward 2.87 -> 2.59 ms per 124 steps; AGNN_GRU_PDL=0
// turns it off).  A waiting grid must not sit BESIDE the running one: the forward kernel is one CTA per SM by registers;
// the backward kernel, of which two fit, asks for more than half an SM of shared memory for that reason -- without
// the padding the waiting grids piled up on the SMs of the running one and the loop got 9x slower (2.9 -> 27 ms).
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("AGNN_GRU_PDL"); return !(e && *e == '0'); }();
  return on;
}
int launch_steps(void (*kernel)(GruParams, int), dim3 grid, size_t smem, const GruParams& p, cudaStream_t st) {
  const bool pdl = pdl_enabled();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  for (int step = 0; step < p.steps; ++step) {
    // step 0 keeps the full stream dependency: its prologue reads what the kernels BEFORE the time loop wrote (gi, dout)
    cfg.numAttrs = pdl && step > 0 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p, step);
    if (e != cudaSuccess) return fail(AGNN_ERR_CUDA, "gru time step %d: %s", step, cudaGetErrorString(e));
  }
  return AGNN_OK;
}

int launch_fwd(const GruParams& p, cudaStream_t st) {
  constexpr size_t kSmem = smem_bytes<FTS, FTN>();
  static const cudaError_t attr =
      cudaFuncSetAttribute(gru_wide_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
  if (attr != cudaSuccess) return fail(AGNN_ERR_CUDA, "gru_fwd: shared-memory attribute: %s", cudaGetErrorString(attr));
  dim3 grid((unsigned)(p.hidden / 8), (unsigned)ceil_div(p.batch, FTS), (unsigned)p.n_dir);
  int rc = launch_steps(gru_wide_fwd_kernel, grid, kSmem, p, st);
  if (rc) return rc;
  return check_launch("gru_wide_fwd");
}

int launch_bwd(const GruParams& p, cudaStream_t st) {
  const size_t kSmem = pdl_enabled() ? (size_t)120 * 1024 : smem_bytes<BTS, BTN>();   // one CTA per SM: see launch_steps
  static const cudaError_t attr =
      cudaFuncSetAttribute(gru_wide_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
  if (attr != cudaSuccess) return fail(AGNN_ERR_CUDA, "gru_bwd: shared-memory attribute: %s", cudaGetErrorString(attr));
  dim3 grid((unsigned)(p.hidden / BTN), (unsigned)ceil_div(p.batch, BTS), (unsigned)p.n_dir);
  int rc = launch_steps(gru_wide_bwd_kernel, grid, kSmem, p, st);
  if (rc) return rc;
  return check_launch("gru_wide_bwd");
}
}  // namespace wide

// ---------------------------------------------------------------------------------------------------------
// Hidden size 128 on the tensor cores, 8 sequences per CTA.  The register-resident kernels above give every pair of
// sequences an SM of its own: 100 of the 148 SMs for the 0.65 - 0.9 ms of each of the four recurrence launches of a
// step, during which the kernels of the main stream (persistent GEMMs, gathers) run on what is left.  Here a CTA owns
// 8 sequences of one direction -- one n8 tile -- so 100 sequences x 2 directions are 26 CTAs, and the recurrent product
//     gh^T[384, 8] = W_hh . h^T[128, 8]          (forward)        dh^T[128, 8] = W_hh^T . dgh^T[384, 8]   (backward)
// runs as mma.sync m16n8k16 on fp16 hi / lo pairs (3 MMAs per product, fp32-grade: the F16X3 form of gemm.cu):
//   * W_hh is split ONCE per launch with one power-of-two scale (its amax, reduced by the CTA) and kept as A
//     FRAGMENTS: three quarters of every warp's fragments in registers (144), the rest in shared memory in fragment
//     order (two LDS.128 per MMA triple) -- 49 152 words in total, which the registers alone do not hold;
//   * h is bounded by 1 (scale 2^13), dgh gets a per-step scale from the CTA's maximum; the B operand lives in shared
//     memory as two [8][128 + 8] fp16 matrices, double buffered, so a time step needs ONE barrier (backward: two);
//   * warp w owns hidden units [16 w, 16 w + 16): its three (forward) / one (backward) m-tiles hold exactly the r, z, n
//     pre-activations / the dh of the units whose gates it evaluates, so the accumulators never leave the thread:
//     lane (g, t) finishes units 16 w + g, 16 w + g + 8 of sequences 2 t, 2 t + 1 and carries their h / dh in
//     registers at full precision;
//   * the per-step inputs of those 4 (sequence, unit) items are requested straight into registers before the MMA phase
//     (8 lanes = one 32-byte sector) and consumed after it: no staging ring.
// (A first version had the 16 sequences of a CTA on the M side: twice the MMAs per CTA and step, 2.9 us per step, and
// with 14 CTAs the loop itself became the critical path of the training step: 9.6 -> 10.4 ms.)
// An accumulation chain is 8 k steps x 3 = 24 MMAs (forward), the limit gemm.cu keeps for the truncating accumulator;
// the backward's 72 are cut into three chains summed with fp32 adds.
namespace tc {
constexpr int H = 128;
constexpr int kThreads = 256;
constexpr int kSeqs = 8;              // one n8 tile
constexpr int LDH = H + 8;            // halves per row of the forward A operand: 68 words = 4 banks per row
constexpr int LDG = 3 * H + 8;        // backward A operand (dgh): 196 words = 4 banks per row
constexpr float kHScale = 8192.f;     // |h| <= 1 -> [0, 2^13]

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Gate nonlinearities on the special-function unit: ex2.approx + rcp.approx (2 ulp each) instead of libm's expf / tanhf
// sequences -- the gate phase of a step is latency bound with two warps per scheduler, and these were 40 % of its
// instructions.  Absolute error ~2e-7 on values in [-1, 1] (parity tests unchanged).
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// Four 8 x 8 fp16 matrices = the B fragments {b0, b1} of TWO k steps from an operand stored [n][k]: lane i passes the
// address of row (i & 7), k offset 8 (i >> 3); lane (g, t) receives the words (row g, k 2t..2t+1) of each matrix.
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(smem_row)));
}

// (x0, x1) * s -> packed fp16 hi pair and lo pair
__device__ __forceinline__ void split_pair(float x0, float x1, float s, uint32_t& hi, uint32_t& lo) {
  const float y0 = x0 * s, y1 = x1 * s;
  const __half2 h = __floats2half2_rn(y0, y1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// max |w| over a [rows x 128] matrix, by the whole CTA (red: 8 floats of shared memory); ends with a barrier
__device__ __forceinline__ float cta_amax(const float* w, int n4, float* red) {
  float m = 0.f;
  for (int i = threadIdx.x; i < n4; i += kThreads) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < kThreads / 32; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  return m;
}

// ---- forward -------------------------------------------------------------------------------------------
// gh^T[384, 8] = W_hh[384, 128] . h^T[128, 8]: the gate rows are the M side (24 m16 tiles, three per warp: the r, z, n
// rows of its 16 units), the 8 sequences one n8 tile: 3 x 8 x 3 = 72 MMAs per warp and step, none wasted on padding.
constexpr int FKS = H / 16;           // 8 k steps
constexpr int FKREG = 6;              // k steps whose W_hh fragments stay in registers (6 x 3 gates x 8 words = 144)
struct FwdSmem {
  uint4 wfrag[kThreads / 32][3][FKS - FKREG][2][32];   // {a0, a1, a2, a3} hi / lo per (warp, gate, k step, lane)
  __half h_hi[2][kSeqs][LDH], h_lo[2][kSeqs][LDH];     // B operand: h as [sequence][unit], double buffered
  float bias[3 * H];
  float red[kThreads / 32];
};

__global__ void __launch_bounds__(kThreads, 1) gru_tc_fwd_kernel(const __grid_constant__ GruParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int dir = blockIdx.y, b0 = blockIdx.x * kSeqs, T = p.steps;
  const float* W = p.w_hh[dir];
  const float s_w = f16_scale_of(cta_amax(W, 3 * H * H / 4, sm.red));
  const float inv = 1.f / (s_w * kHScale);
  // A fragments (row, k) of gate q, k step ks: (g, 2t..) (g + 8, 2t..) (g, 2t + 8..) (g + 8, 2t + 8..), rows q H + 16 w + .
  uint32_t wreg[3][FKREG][8];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float* wr = W + (int64_t)(q * H + 16 * w + g) * H + 2 * t;
#pragma unroll
    for (int ks = 0; ks < FKS; ++ks) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(wr + (i & 1) * 8 * H + 16 * ks + (i >> 1) * 8));
        split_pair(v.x, v.y, s_w, hi[i], lo[i]);
      }
      if (ks < FKREG) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { wreg[q][ks][i] = hi[i]; wreg[q][ks][4 + i] = lo[i]; }
      } else {
        sm.wfrag[w][q][ks - FKREG][0][lane] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        sm.wfrag[w][q][ks - FKREG][1][lane] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
  for (int i = tid; i < 3 * H; i += kThreads) sm.bias[i] = __ldg(p.b_hh[dir] + i);
  for (int i = tid; i < 2 * kSeqs * LDH / 2; i += kThreads) {          // h_0 = 0 in both buffers
    reinterpret_cast<uint32_t*>(&sm.h_hi[0][0][0])[i] = 0u;
    reinterpret_cast<uint32_t*>(&sm.h_lo[0][0][0])[i] = 0u;
  }
  __syncthreads();
  // this thread finishes units 16 w + g + 8 s2 (s2 = 0, 1) of sequences 2 t + e (e = 0, 1): C fragment [2 s2 + e]
  float hprev[2][2], bias[3][2];
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    hprev[s2][0] = hprev[s2][1] = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q) bias[q][s2] = sm.bias[q * H + 16 * w + g + 8 * s2];
  }
  // Addresses: fixed offsets from three pointers that move by one time step per iteration
  const int dsign = dir == 0 ? 1 : -1;
  const int64_t row0 = (int64_t)(b0 + 2 * t) * T + (dir == 0 ? 0 : T - 1);
  const int64_t so_gi = (int64_t)T * 3 * H, so_out = (int64_t)T * p.ld_out, so_gt = (int64_t)T * 4 * H;   // sequence e = 1
  const float* gi_p = p.gi[dir] + row0 * (3 * H) + 16 * w + g;
  float* out_p = p.out + row0 * p.ld_out + dir * H + 16 * w + g;
  float* gt_p = p.gates[dir] ? p.gates[dir] + row0 * (4 * H) + 16 * w + g : nullptr;
  const bool ok[2] = {b0 + 2 * t < p.batch, b0 + 2 * t + 1 < p.batch};
  int cur = 0;
  for (int step = 0; step < T; ++step) {
    // this thread's input projections of the step, requested now and used after the MMA phase
    float gi[3][2][2];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
        for (int e = 0; e < 2; ++e) gi[q][s2][e] = ok[e] ? __ldg(gi_p + e * so_gi + q * H + 8 * s2) : 0.f;
    float acc[3][4];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[q][i] = 0.f;
    if (step > 0) {                                                     // h_0 = 0: nothing to multiply
      // B fragments (k, n): (2t.., g) (2t + 8.., g) = h[sequence g][units 16 ks + 2t.. / + 8], two k steps per ldmatrix
      const __half* bh_row = &sm.h_hi[cur][lane & 7][8 * (lane >> 3)];
      const __half* bl_row = &sm.h_lo[cur][lane & 7][8 * (lane >> 3)];
      uint32_t bh[4], bl[4];
#pragma unroll
      for (int ks = 0; ks < FKS; ++ks) {
        if ((ks & 1) == 0) {
          ldmatrix_x4(bh, bh_row + 16 * ks);
          ldmatrix_x4(bl, bl_row + 16 * ks);
        }
        const uint32_t bh0 = bh[2 * (ks & 1)], bh1 = bh[2 * (ks & 1) + 1], bl0 = bl[2 * (ks & 1)], bl1 = bl[2 * (ks & 1) + 1];
        uint32_t ah[3][4], al[3][4];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (ks < FKREG) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { ah[q][i] = wreg[q][ks][i]; al[q][i] = wreg[q][ks][4 + i]; }
          } else {
            const uint4 vh = sm.wfrag[w][q][ks - FKREG][0][lane], vl = sm.wfrag[w][q][ks - FKREG][1][lane];
            ah[q][0] = vh.x; ah[q][1] = vh.y; ah[q][2] = vh.z; ah[q][3] = vh.w;
            al[q][0] = vl.x; al[q][1] = vl.y; al[q][2] = vl.z; al[q][3] = vl.w;
          }
        }
        // term-major: consecutive MMAs go to different accumulators
#pragma unroll
        for (int q = 0; q < 3; ++q) mma_f16(acc[q], al[q], bh0, bh1);   // W_lo h_hi
#pragma unroll
        for (int q = 0; q < 3; ++q) mma_f16(acc[q], ah[q], bl0, bl1);   // W_hi h_lo
#pragma unroll
        for (int q = 0; q < 3; ++q) mma_f16(acc[q], ah[q], bh0, bh1);   // W_hi h_hi
      }
    }
    const int nxt = cur ^ 1;
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const int u = 16 * w + g + 8 * s2;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float ghr = acc[0][2 * s2 + e] * inv + bias[0][s2];
        const float ghz = acc[1][2 * s2 + e] * inv + bias[1][s2];
        const float ghn = acc[2][2 * s2 + e] * inv + bias[2][s2];
        const float r = fast_sigmoid(gi[0][s2][e] + ghr);
        const float z = fast_sigmoid(gi[1][s2][e] + ghz);
        const float n = fast_tanh(gi[2][s2][e] + r * ghn);
        const float hn = (1.f - z) * n + z * hprev[s2][e];
        hprev[s2][e] = hn;
        const float y = hn * kHScale;
        const __half hh = __float2half_rn(y);
        sm.h_hi[nxt][2 * t + e][u] = hh;
        sm.h_lo[nxt][2 * t + e][u] = __float2half_rn(y - __half2float(hh));
        if (ok[e]) {
          out_p[e * so_out + 8 * s2] = hn;
          if (gt_p) {
            float* gt = gt_p + e * so_gt + 8 * s2;
            gt[0] = r; gt[H] = z; gt[2 * H] = n; gt[3 * H] = ghn;
          }
        }
      }
    }
    __syncthreads();                  // the new operand is complete; the old one is no longer read
    cur = nxt;
    gi_p += dsign * 3 * H; out_p += (int64_t)dsign * p.ld_out;
    if (gt_p) gt_p += dsign * 4 * H;
  }
}

// ---- backward ------------------------------------------------------------------------------------------
// dh^T[128, 8] = W_hh^T[128, 384] . dgh^T[384, 8]: warp w owns the m16 tile of its 16 units, 24 k steps.  The three
// product terms accumulate in three separate fragments (independent chains of 24 MMAs each, summed small-first with
// fp32 adds at the end): one fragment would make the 72 MMAs of a step wait for each other.
constexpr int BKS = 3 * H / 16;       // 24 k steps
constexpr int BKREG = 16;             // k steps whose W_hh^T fragments stay in registers (16 x 8 words = 128)
struct BwdSmem {
  uint4 wfrag[kThreads / 32][BKS - BKREG][2][32];      // {a0, a1, a2, a3} hi / lo per (warp, k step, lane)
  __half d_hi[2][kSeqs][LDG], d_lo[2][kSeqs][LDG];     // B operand: dgh as [sequence][gate index], double buffered
  float red[2][kThreads / 32];                         // per-step maxima of the warps, double buffered
};

__global__ void __launch_bounds__(kThreads, 1) gru_tc_bwd_kernel(const __grid_constant__ GruParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int dir = blockIdx.y, b0 = blockIdx.x * kSeqs, T = p.steps;
  const float* W = p.w_hh[dir];
  const float s_w = f16_scale_of(cta_amax(W, 3 * H * H / 4, sm.red[0]));
  // A = W_hh^T: A[u][i] = W_hh[i][u]; fragment (row, k): (g, 2t..) (g + 8, 2t..) (g, 2t + 8..) (g + 8, 2t + 8..)
  uint32_t wreg[BKREG][8];
#pragma unroll
  for (int ks = 0; ks < BKS; ++ks) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* wp = W + (int64_t)(16 * ks + 2 * t + (i >> 1) * 8) * H + 16 * w + g + (i & 1) * 8;
      split_pair(__ldg(wp), __ldg(wp + H), s_w, hi[i], lo[i]);
    }
    if (ks < BKREG) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { wreg[ks][i] = hi[i]; wreg[ks][4 + i] = lo[i]; }
    } else {
      sm.wfrag[w][ks - BKREG][0][lane] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      sm.wfrag[w][ks - BKREG][1][lane] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  __syncthreads();
  // this thread finishes units 16 w + g + 8 s2 (s2 = 0, 1) of sequences 2 t + e (e = 0, 1): C fragment [2 s2 + e]
  float keep[2][2] = {{0.f, 0.f}, {0.f, 0.f}};          // z * dh of the step before: the direct path to h_prev
  float inv = 0.f;                                       // 1 / (s_w * scale of the dgh operand in sm.d_*[cur])
  float gmax = 0.f;
  const bool ok[2] = {b0 + 2 * t < p.batch, b0 + 2 * t + 1 < p.batch};
  // Addresses: everything this thread touches in a step sits at fixed offsets from five pointers that move by one
  // time step per iteration (the forward's order reversed: t = T-1 .. 0 for the forward direction, 0 .. T-1 for the
  // reverse one); h_prev of the forward recurrence is one more step along the same way.
  const int dsign = dir == 0 ? -1 : 1;
  const int64_t row0 = (int64_t)(b0 + 2 * t) * T + (dir == 0 ? T - 1 : 0);
  const int64_t so_out = (int64_t)T * p.ld_out, so_gt = (int64_t)T * 4 * H, so_g3 = (int64_t)T * 3 * H;   // sequence e = 1
  const float* dout_p = p.dout + row0 * p.ld_out + dir * H + 16 * w + g;
  const float* out_p = p.out + row0 * p.ld_out + dir * H + 16 * w + g;
  const float* gt_p = p.gates[dir] + row0 * (4 * H) + 16 * w + g;
  float* dgi_p = p.dgi[dir] + row0 * (3 * H) + 16 * w + g;
  float* dgh_p = p.dgh[dir] + row0 * (3 * H) + 16 * w + g;
  const int64_t d_out = (int64_t)dsign * p.ld_out;
  int cur = 0;
  for (int step = 0; step < T; ++step) {
    // this thread's inputs of the step, requested now and used after the MMA phase
    float gd[2][2], r[2][2], z[2][2], n[2][2], ghn[2][2], hp[2][2];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        gd[s2][e] = r[s2][e] = z[s2][e] = n[s2][e] = ghn[s2][e] = hp[s2][e] = 0.f;
        if (ok[e]) {
          const float* gt = gt_p + e * so_gt + 8 * s2;
          gd[s2][e] = __ldg(dout_p + e * so_out + 8 * s2);
          r[s2][e] = __ldg(gt); z[s2][e] = __ldg(gt + H); n[s2][e] = __ldg(gt + 2 * H); ghn[s2][e] = __ldg(gt + 3 * H);
          if (step < T - 1) hp[s2][e] = __ldg(out_p + d_out + e * so_out + 8 * s2);
        }
      }
    float acc[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
    if (step > 0) {
      // B fragments (k, n): (2t.., g) (2t + 8.., g) = dgh[sequence g][gate index 16 ks + 2t.. / + 8]
      const __half* bh_row = &sm.d_hi[cur][lane & 7][8 * (lane >> 3)];
      const __half* bl_row = &sm.d_lo[cur][lane & 7][8 * (lane >> 3)];
      uint32_t bh[4], bl[4];
#pragma unroll
      for (int ks = 0; ks < BKS; ++ks) {
        if ((ks & 1) == 0) {
          ldmatrix_x4(bh, bh_row + 16 * ks);
          ldmatrix_x4(bl, bl_row + 16 * ks);
        }
        const uint32_t bh0 = bh[2 * (ks & 1)], bh1 = bh[2 * (ks & 1) + 1], bl0 = bl[2 * (ks & 1)], bl1 = bl[2 * (ks & 1) + 1];
        uint32_t ah[4], al[4];
        if (ks < BKREG) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { ah[i] = wreg[ks][i]; al[i] = wreg[ks][4 + i]; }
        } else {
          const uint4 vh = sm.wfrag[w][ks - BKREG][0][lane], vl = sm.wfrag[w][ks - BKREG][1][lane];
          ah[0] = vh.x; ah[1] = vh.y; ah[2] = vh.z; ah[3] = vh.w;
          al[0] = vl.x; al[1] = vl.y; al[2] = vl.z; al[3] = vl.w;
        }
        mma_f16(acc[0], al, bh0, bh1);                    // W_lo dgh_hi
        mma_f16(acc[1], ah, bl0, bl1);                    // W_hi dgh_lo
        mma_f16(acc[2], ah, bh0, bh1);                    // W_hi dgh_hi
      }
    }
    float dr[2][2], dz[2][2], dn[2][2], dgn[2][2];
    float smax = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 2 * s2 + e;
        const float mat = ((acc[0][i] + acc[1][i]) + acc[2][i]) * inv;
        const float dh = gd[s2][e] + (keep[s2][e] + mat);
        dn[s2][e] = dh * (1.f - z[s2][e]) * (1.f - n[s2][e] * n[s2][e]);
        dz[s2][e] = dh * (hp[s2][e] - n[s2][e]) * z[s2][e] * (1.f - z[s2][e]);
        dr[s2][e] = dn[s2][e] * ghn[s2][e] * r[s2][e] * (1.f - r[s2][e]);
        dgn[s2][e] = dn[s2][e] * r[s2][e];
        keep[s2][e] = dh * z[s2][e];
        smax = fmaxf(smax, fmaxf(fabsf(dr[s2][e]), fmaxf(fabsf(dz[s2][e]), fabsf(dn[s2][e]))));   // >= |dgn|
        if (ok[e]) {
          float* gi_o = dgi_p + e * so_g3 + 8 * s2;
          float* gh_o = dgh_p + e * so_g3 + 8 * s2;
          gi_o[0] = dr[s2][e]; gi_o[H] = dz[s2][e]; gi_o[2 * H] = dn[s2][e];
          gh_o[0] = dr[s2][e]; gh_o[H] = dz[s2][e]; gh_o[2 * H] = dgn[s2][e];
        }
      }
    gmax = fmaxf(gmax, smax);
    // the operand scale of this step's dgh: the CTA's maximum (padding sequences contribute zeros)
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, d));
    if (lane == 0) sm.red[step & 1][w] = smax;
    __syncthreads();
    float m = sm.red[step & 1][0];
#pragma unroll
    for (int i = 1; i < kThreads / 32; ++i) m = fmaxf(m, sm.red[step & 1][i]);
    const float s_d = f16_scale_of(m);
    inv = 1.f / (s_w * s_d);
    const int nxt = cur ^ 1;
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int u = 16 * w + g + 8 * s2;
        const float v[3] = {dr[s2][e], dz[s2][e], dgn[s2][e]};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const float y = v[q] * s_d;
          const __half hh = __float2half_rn(y);
          sm.d_hi[nxt][2 * t + e][q * H + u] = hh;
          sm.d_lo[nxt][2 * t + e][q * H + u] = __float2half_rn(y - __half2float(hh));
        }
      }
    __syncthreads();                  // the new operand is complete; the old one and red[step & 1] are no longer read
    cur = nxt;
    dout_p += d_out; out_p += d_out;
    gt_p += dsign * 4 * H; dgi_p += dsign * 3 * H; dgh_p += dsign * 3 * H;
  }
  if (p.amax[dir]) {
    uint32_t m = __float_as_uint(gmax);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0 && m) atomicMax(reinterpret_cast<unsigned int*>(p.amax[dir]), m);
  }
}

// Which recurrence kernels hidden size 128 takes (agnn_gru_mode; initial value from AGNN_GRU_TC): 0 = register-resident
// SIMT kernels for both passes, 1 = both passes on the tensor cores, 2 = forward only (default), 3 = backward only.
// Measured on the headline training step (B200, profiles/README.md): 0: 9.53 ms, 2: 9.19 ms, 1: 9.63 ms -- the forward
// loop is 0.82 ms on 26 SMs instead of 0.65 ms on 100 and the main stream gets the difference; the backward loop
// (1.22 ms on 26 SMs against 0.875 ms on 100: a second barrier per step for the operand scale, 632 instead of 443
// instructions per warp and step) ends later than the main stream's backward and stays on the SIMT kernel.
inline int& mode() {
  static int m = [] { const char* e = getenv("AGNN_GRU_TC"); return e && *e ? atoi(e) : 2; }();
  return m;
}
// mode 2 takes the tensor-core forward only when the SIMT kernel would hold more than half of the SMs (one CTA per 2
// sequences and direction): below that the SM-time it frees is small and its longer loop (0.76 vs 0.65 ms) is the
// critical path of a short step -- strong scaling of the headline batch over 2 GPUs (50 sequences per rank): 6.25 ms
// with the SIMT forward, 6.52 ms with the tensor-core one.
inline bool fwd_enabled(int batch, int n_dir) {
  return mode() == 1 || (mode() == 2 && ceil_div(batch, kSeq) * n_dir > kNumSM / 2);
}
inline bool bwd_enabled() { return mode() == 1 || mode() == 3; }

int launch_fwd(const GruParams& p, cudaStream_t st) {
  static const cudaError_t attr =
      cudaFuncSetAttribute(gru_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem));
  if (attr != cudaSuccess) return fail(AGNN_ERR_CUDA, "gru_fwd: shared-memory attribute: %s", cudaGetErrorString(attr));
  dim3 grid((unsigned)ceil_div(p.batch, kSeqs), (unsigned)p.n_dir);
  gru_tc_fwd_kernel<<<grid, kThreads, sizeof(FwdSmem), st>>>(p);
  return check_launch("gru_tc_fwd");
}

int launch_bwd(const GruParams& p, cudaStream_t st) {
  static const cudaError_t attr =
      cudaFuncSetAttribute(gru_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem));
  if (attr != cudaSuccess) return fail(AGNN_ERR_CUDA, "gru_bwd: shared-memory attribute: %s", cudaGetErrorString(attr));
  dim3 grid((unsigned)ceil_div(p.batch, kSeqs), (unsigned)p.n_dir);
  gru_tc_bwd_kernel<<<grid, kThreads, sizeof(BwdSmem), st>>>(p);
  return check_launch("gru_tc_bwd");
}
}  // namespace tc

int check_gru(const char* what, int batch, int steps, int hidden, int n_dir) {
  if (batch < 0 || steps < 0 || (n_dir != 1 && n_dir != 2))
    return fail(AGNN_ERR_ARG, "%s: bad sizes (batch=%d steps=%d n_dir=%d)", what, batch, steps, n_dir);
  if (hidden != 32 && hidden != 64 && hidden != 128 && !wide::supported(hidden))
    return fail(AGNN_ERR_UNSUPPORTED, "%s: hidden size %d (register-resident W_hh: 32, 64, 128; per-step launches: "
                "multiples of 64 up to 2048)", what, hidden);
  return AGNN_OK;
}

template <int H>
int launch_fwd(const GruParams& p, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(p.batch, kSeq), (unsigned)p.n_dir);
  gru_fwd_kernel<H><<<grid, 3 * H, 0, st>>>(p);
  return check_launch("gru_fwd");
}
template <int H>
int launch_bwd(const GruParams& p, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(p.batch, kSeq), (unsigned)p.n_dir);
  gru_bwd_kernel<H><<<grid, 3 * H, 0, st>>>(p);
  return check_launch("gru_bwd");
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_gru_mode(int mode) {
  const int old = tc::mode();
  if (mode >= 0 && mode <= 3) tc::mode() = mode;
  return old;
}

extern "C" int agnn_gru_supported(int hidden) {
  if (hidden == 32 || hidden == 64 || hidden == 128) return AGNN_GRU_RESIDENT;
  return wide::supported(hidden) ? AGNN_GRU_STEPWISE : 0;
}

extern "C" int agnn_gru_fwd(int32_t batch, int32_t steps, int32_t hidden, int32_t n_dir, const float* const* gi,
                            const float* const* w_hh, const float* const* b_hh, float* out, float* const* gates,
                            agnn_stream_t stream) {
  int rc = check_gru("gru_fwd", batch, steps, hidden, n_dir);
  if (rc) return rc;
  if (!gi || !w_hh || !b_hh || !out) return fail(AGNN_ERR_ARG, "gru_fwd: null pointer");
  if (batch == 0 || steps == 0) return AGNN_OK;
  GruParams p;
  memset(&p, 0, sizeof(p));
  p.batch = batch; p.steps = steps; p.n_dir = n_dir; p.out = out; p.ld_out = (int64_t)n_dir * hidden;
  for (int d = 0; d < n_dir; ++d) {
    if (!gi[d] || !w_hh[d] || !b_hh[d] || !aligned16(w_hh[d])) return fail(AGNN_ERR_ARG, "gru_fwd: null / unaligned operand");
    p.gi[d] = gi[d]; p.w_hh[d] = w_hh[d]; p.b_hh[d] = b_hh[d];
    p.gates[d] = gates ? gates[d] : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  p.hidden = hidden;
  if (wide::supported(hidden)) {
    for (int d = 0; d < n_dir; ++d)
      if (!aligned16(gi[d]) || !aligned16(out) || !aligned16(b_hh[d]))
        return fail(AGNN_ERR_ARG, "gru_fwd: unaligned operand");
    return wide::launch_fwd(p, st);
  }
  if (hidden == 128 && tc::fwd_enabled(batch, n_dir)) {
    for (int d = 0; d < n_dir; ++d)
      if (!aligned16(gi[d]) || !aligned16(out) || (gates && gates[d] && !aligned16(gates[d])))
        return fail(AGNN_ERR_ARG, "gru_fwd: unaligned operand");
    return tc::launch_fwd(p, st);
  }
  if (hidden == 128) return launch_fwd<128>(p, st);
  if (hidden == 64) return launch_fwd<64>(p, st);
  return launch_fwd<32>(p, st);
}

extern "C" int agnn_gru_bwd(int32_t batch, int32_t steps, int32_t hidden, int32_t n_dir, const float* const* w_hh,
                            const float* out, const float* const* gates, const float* dout, float* const* dgi,
                            float* const* dgh, agnn_stream_t stream) {
  return agnn_gru_bwd_amax(batch, steps, hidden, n_dir, w_hh, out, gates, dout, dgi, dgh, nullptr, stream);
}

extern "C" int agnn_gru_bwd_amax(int32_t batch, int32_t steps, int32_t hidden, int32_t n_dir, const float* const* w_hh,
                                 const float* out, const float* const* gates, const float* dout, float* const* dgi,
                                 float* const* dgh, float* const* amax, agnn_stream_t stream) {
  int rc = check_gru("gru_bwd", batch, steps, hidden, n_dir);
  if (rc) return rc;
  if (wide::supported(hidden))
    return fail(AGNN_ERR_UNSUPPORTED, "gru_bwd: hidden size %d takes agnn_gru_bwd_stepwise (W_hh^T + carry buffer)", hidden);
  if (!w_hh || !out || !gates || !dout || !dgi || !dgh) return fail(AGNN_ERR_ARG, "gru_bwd: null pointer");
  if (batch == 0 || steps == 0) return AGNN_OK;
  GruParams p;
  memset(&p, 0, sizeof(p));
  p.batch = batch; p.steps = steps; p.n_dir = n_dir; p.out = const_cast<float*>(out);
  p.ld_out = (int64_t)n_dir * hidden; p.dout = dout;
  for (int d = 0; d < n_dir; ++d) {
    if (!w_hh[d] || !gates[d] || !dgi[d] || !dgh[d]) return fail(AGNN_ERR_ARG, "gru_bwd: null operand");
    p.w_hh[d] = w_hh[d]; p.gates[d] = const_cast<float*>(gates[d]); p.dgi[d] = dgi[d]; p.dgh[d] = dgh[d];
    p.amax[d] = amax ? amax[d] : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (hidden == 128 && tc::bwd_enabled()) return tc::launch_bwd(p, st);
  if (hidden == 128) return launch_bwd<128>(p, st);
  if (hidden == 64) return launch_bwd<64>(p, st);
  return launch_bwd<32>(p, st);
}

extern "C" int agnn_gru_bwd_stepwise(int32_t batch, int32_t steps, int32_t hidden, int32_t n_dir,
                                     const float* const* w_hh_t, const float* out, const float* const* gates,
                                     const float* dout, float* const* dgi, float* const* dgh, float* const* amax,
                                     float* carry, agnn_stream_t stream) {
  int rc = check_gru("gru_bwd_stepwise", batch, steps, hidden, n_dir);
  if (rc) return rc;
  if (!wide::supported(hidden))
    return fail(AGNN_ERR_UNSUPPORTED, "gru_bwd_stepwise: hidden size %d takes agnn_gru_bwd", hidden);
  if (!w_hh_t || !out || !gates || !dout || !dgi || !dgh || !carry) return fail(AGNN_ERR_ARG, "gru_bwd_stepwise: null pointer");
  if (batch == 0 || steps == 0) return AGNN_OK;
  GruParams p;
  memset(&p, 0, sizeof(p));
  p.batch = batch; p.steps = steps; p.n_dir = n_dir; p.hidden = hidden; p.out = const_cast<float*>(out);
  p.ld_out = (int64_t)n_dir * hidden; p.dout = dout; p.keep = carry;
  if (!aligned16(out) || !aligned16(dout) || !aligned16(carry)) return fail(AGNN_ERR_ARG, "gru_bwd_stepwise: unaligned operand");
  for (int d = 0; d < n_dir; ++d) {
    if (!w_hh_t[d] || !gates[d] || !dgi[d] || !dgh[d] || !aligned16(w_hh_t[d]) || !aligned16(dgh[d]) ||
        !aligned16(dgi[d]) || !aligned16(gates[d]))
      return fail(AGNN_ERR_ARG, "gru_bwd_stepwise: null / unaligned operand");
    p.w_hh_t[d] = w_hh_t[d]; p.gates[d] = const_cast<float*>(gates[d]); p.dgi[d] = dgi[d]; p.dgh[d] = dgh[d];
    p.amax[d] = amax ? amax[d] : nullptr;
  }
  return wide::launch_bwd(p, (cudaStream_t)stream);
}
