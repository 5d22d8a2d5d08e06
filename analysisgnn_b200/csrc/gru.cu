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

int check_gru(const char* what, int batch, int steps, int hidden, int n_dir) {
  if (batch < 0 || steps < 0 || (n_dir != 1 && n_dir != 2))
    return fail(AGNN_ERR_ARG, "%s: bad sizes (batch=%d steps=%d n_dir=%d)", what, batch, steps, n_dir);
  if (hidden != 32 && hidden != 64 && hidden != 128)
    return fail(AGNN_ERR_UNSUPPORTED, "%s: hidden size %d (register-resident W_hh supports 32, 64, 128)", what, hidden);
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

extern "C" int agnn_gru_supported(int hidden) { return hidden == 32 || hidden == 64 || hidden == 128; }

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
  if (hidden == 128) return launch_bwd<128>(p, st);
  if (hidden == 64) return launch_bwd<64>(p, st);
  return launch_bwd<32>(p, st);
}
