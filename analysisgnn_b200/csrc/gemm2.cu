// 2-CTA (cta_group::2) variant of the fp32-parity projection GEMM for the LARGE products of a step
// (50 000 x 256 x 2560 and transposes: the fused message-passing layers' forward and grad-input products).
//
// Why: the single-CTA kernel (gemm.cu) moves 64 KB of operand tiles into shared memory per 128 x 128 x 64 K block
// (A hi/lo + B hi/lo); on the big shapes that is 5.9 KB / clk chip wide -- the L2 -> SM fill rate, not the tensor pipe
// (64-70 % active), bounds them.  A CTA pair computing a 256 x 256 tile with tcgen05.mma.cta_group::2 keeps the same
// 64 KB per CTA and K block (its 128 rows of A, its 128-column HALF of B) for twice the MMA work: half the fill
// traffic per flop.
//
// Structure (per CTA of the pair): warp 0 = TMA producer (own A rows, own half of B, own barriers), warp 1 = MMA
// issuer in the leader CTA / stage forwarder in the peer (tells the leader that the peer's stage has landed),
// warps 2..9 = epilogue (8 warps: TMEM lane quarter x column half), same chained accumulation as gemm.cu: a TMEM
// chain holds 2 K blocks (24 MMAs), the epilogue warps add the chains in fp32 registers with round-to-nearest.
// Synchronisation: full[s] local per CTA; peer_full[s] in the leader (remote arrive by the peer's forwarder);
// empty[s] / acc_full[b] in BOTH CTAs (tcgen05.commit ... multicast::cluster); acc_empty[b] in the leader only
// (16 arrivals: 8 local + 8 remote epilogue warps).  F16X3 operands (fp16 hi / lo pairs with per-tensor scales).
#include <cuda.h>

#include <cstring>

#include "common.cuh"

namespace agnn {
namespace {

constexpr int kBlockM = 128;                // rows per CTA (256 per pair)
constexpr int kBlockN = 256;                // columns per pair tile; a CTA stages 128 of them
constexpr int kHalfN = 128;
constexpr int kRowBytes = 128;
constexpr int kTileBytes = 128 * kRowBytes; // 16 KB: one operand tile (hi or lo) of one CTA
constexpr int kBlockK = 64;                 // fp16 elements per 128-byte row
constexpr int kUmmaK = 16;
constexpr int kStageBytes = 4 * kTileBytes; // A hi, A lo, B hi, B lo
constexpr int kStages = 3;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kTmemCols = 512;              // 2 accumulator buffers x 256 columns
constexpr int kStoreBox = 32;
constexpr int kStoreBufBytes = kStoreBox * 128;
constexpr int kStoreBytes = kEpiWarps * kStoreBufBytes;   // one staging box per epilogue warp
constexpr int kChunk = 64, kChunks = 2;     // MN-major tiles: 64-element chunks of the 128 MN extent

struct Gemm2Params {
  CUtensorMap map_a[2], map_b[2], map_c;
  int M, N, K, k_blocks, tiles_m, tiles_n, chain_blocks;
  const float* bias;
  int flags;
  const float* amax_a;
  const float* amax_b;
  float* amax_out;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// wait with cluster-scope acquire: the arrivals come from the other CTA (or from tcgen05.commit multicasts)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "DONE_C:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MMAs issued so far have completed -> arrive on `bar` (same offset) in both CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = fp32, A / B = fp16, M = 256 across the pair
__host__ __device__ constexpr uint32_t instr_desc2(bool b_mn) {
  return (1u << 4) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(kBlockN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) gemm2_kernel(const __grid_constant__ Gemm2Params p) {
  constexpr uint32_t kIdesc = instr_desc2(B_MN);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* store_base = smem + kStages * kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(store_base + kStoreBytes);
  uint64_t* peer_full = full + kStages;     // used in the leader: the peer's stage s has landed
  uint64_t* empty = peer_full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_empty = acc_full + 2;       // used in the leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = rank == 0;
  const int n_tiles = p.tiles_m * p.tiles_n;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      prefetch_tmap(&p.map_a[i]);
      prefetch_tmap(&p.map_b[i]);
    }
    prefetch_tmap(&p.map_c);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&peer_full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 2 * kEpiWarps);   // both CTAs' epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                // barriers of both CTAs are initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs: own rows, own half of B)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        const int m0 = (tile / p.tiles_n) * 256 + (int)rank * kBlockM;
        const int nb = (tile % p.tiles_n) * kBlockN + (int)rank * kHalfN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait_cluster(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * kStageBytes;
          mbar_expect_tx(&full[stage], kStageBytes);
          const int k0 = kb * kBlockK;
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            tma_load_2d(st + part * kTileBytes, &p.map_a[part], &full[stage], k0, m0);
            uint8_t* b_dst = st + (2 + part) * kTileBytes;
            if constexpr (B_MN) {
#pragma unroll
              for (int c = 0; c < kChunks; ++c)
                tma_load_2d(b_dst + c * (kBlockK * kRowBytes), &p.map_b[part], &full[stage], nb + c * kChunk, k0);
            } else {
              tma_load_2d(b_dst, &p.map_b[part], &full[stage], k0, nb);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      if (!leader) {
        // ---------------------------------------------------------- peer: forward "my stage s has landed"
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs)
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&full[stage], phase);
            mbar_arrive_remote(&peer_full[stage], 0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
      } else {
        // ---------------------------------------------------------- leader: MMA issue for the pair
        constexpr uint32_t kHiK = (1024u >> 4) | (1u << 14) | (2u << 29);    // K-major, SWIZZLE_128B
        constexpr uint32_t kHiMn = (1024u >> 4) | (1u << 14) | (2u << 29);   // MN-major 16-bit: SBO 1024, SWIZZLE_128B
        constexpr uint32_t kLboK = (16u >> 4) << 16;
        constexpr uint32_t kLboMn = (((uint32_t)(kBlockK * kRowBytes) >> 4) & 0x3FFF) << 16;
        constexpr uint32_t kStepA = 32 >> 4;
        constexpr uint32_t kStepB = B_MN ? (kUmmaK * kRowBytes) >> 4 : 32 >> 4;
        const uint32_t smem_base = smem_u32(smem);
        int stage = 0;
        uint32_t phase = 0;
        int cc = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs) {
          for (int c0 = 0; c0 < p.k_blocks; c0 += p.chain_blocks, ++cc) {
            const int c1 = min(c0 + p.chain_blocks, p.k_blocks);
            const int buf = cc & 1;
            mbar_wait_cluster(&acc_empty[buf], ((cc >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + buf * kBlockN;
            uint32_t accumulate = 0;
            for (int kb = c0; kb < c1; ++kb) {
              mbar_wait(&full[stage], phase);
              mbar_wait_cluster(&peer_full[stage], phase);
              tc_fence_after();
              const uint32_t st = (smem_base + stage * kStageBytes) >> 4;
#pragma unroll
              for (int term = 0; term < 3; ++term) {   // hi*hi, hi*lo, lo*hi
                const uint32_t a_lo = (st + (term == 2 ? 1 : 0) * (kTileBytes >> 4)) | kLboK;
                const uint32_t b_lo = (st + (2 + (term == 1 ? 1 : 0)) * (kTileBytes >> 4)) | (B_MN ? kLboMn : kLboK);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  const uint64_t da = ((uint64_t)kHiK << 32) | (a_lo + k * kStepA);
                  const uint64_t db = ((uint64_t)(B_MN ? kHiMn : kHiK) << 32) | (b_lo + k * kStepB);
                  umma2(tmem_d, da, db, kIdesc, accumulate);
                  accumulate = 1;
                }
              }
              umma2_commit(&empty[stage]);            // frees the stage in BOTH CTAs
              if (kb == c1 - 1) umma2_commit(&acc_full[buf]);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int ew = warp - 2;
    const int quarter = warp & 3;                 // TMEM lanes [32 * quarter, +32) are this warp's
    const int half = ew >> 2;                     // which 128 of the tile's 256 columns
    uint8_t* sbuf = store_base + ew * kStoreBufBytes;
    const bool scaled = p.amax_a != nullptr;
    const float inv_a = scaled ? 1.f / f16_scale_of(__ldg(p.amax_a)) : 1.f;
    const float inv_b = scaled ? 1.f / f16_scale_of(__ldg(p.amax_b)) : 1.f;
    int cc = 0;
    bool staged = false;
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const int m0 = (tile / p.tiles_n) * 256 + (int)rank * kBlockM;
      const int n0 = (tile % p.tiles_n) * kBlockN + half * kHalfN;
      const int row = m0 + quarter * 32 + lane;
      float acc[kHalfN];
#pragma unroll
      for (int j = 0; j < kHalfN; ++j) acc[j] = 0.f;
      for (int c0 = 0; c0 < p.k_blocks; c0 += p.chain_blocks, ++cc) {
        const int buf = cc & 1;
        mbar_wait_cluster(&acc_full[buf], (cc >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + buf * kBlockN + half * kHalfN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
        for (int c = 0; c < kHalfN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&acc_empty[buf]);
          else mbar_arrive_remote(&acc_empty[buf], 0);
        }
      }
#pragma unroll
      for (int j = 0; j < kHalfN; ++j) {
        float v = (acc[j] * inv_a) * inv_b;
        if (p.bias && n0 + j < p.N) v += __ldg(p.bias + n0 + j);
        if (p.flags & AGNN_GEMM_RELU) v = fmaxf(v, 0.f);
        acc[j] = v;
      }
      if (p.amax_out) {
        uint32_t mx = 0;
        if (row < p.M) {
#pragma unroll
          for (int j = 0; j < kHalfN; ++j)
            if (n0 + j < p.N) mx = max(mx, __float_as_uint(acc[j]) & 0x7fffffffu);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        if (lane == 0 && mx) atomicMax(reinterpret_cast<unsigned int*>(p.amax_out), mx);
      }
#pragma unroll
      for (int c = 0; c < kHalfN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;
        float* v = acc + c * 32;
        if (staged) {                               // one staging box per warp: the previous store must have read it
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        const uint32_t row_addr = smem_u32(sbuf) + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((j ^ (lane & 7)) << 4)),
                       "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                       : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tma_store_2d(&p.map_c, sbuf, col0, m0 + quarter * 32);
        staged = true;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();                                  // nobody leaves while the pair's MMAs may still read its shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn2() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int make_map2(CUtensorMap* map, const void* ptr, bool f16, int64_t rows, int64_t cols, int64_t ld, int box_cols,
              int box_rows) {
  EncodeTiledFn fn = encode_fn2();
  if (!fn) return fail(AGNN_ERR_CUDA, "gemm2: cuTensorMapEncodeTiled is not available from this driver");
  const int eb = f16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * eb};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(AGNN_ERR_CUDA, "gemm2: cuTensorMapEncodeTiled failed (%d)", (int)rc);
  return AGNN_OK;
}

template <bool B_MN>
int launch2(const Gemm2Params& p, int grid, cudaStream_t st) {
  constexpr int smem = kStages * kStageBytes + kStoreBytes + 1024 /*align*/ + 256 /*barriers*/;
  auto kern = gemm2_kernel<B_MN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return check_launch("gemm2: cudaFuncSetAttribute");
    configured = true;
  }
  kern<<<grid, kThreads, smem, st>>>(p);
  return check_launch("gemm2");
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_gemm_pair_supported(int precision, int a_layout, int64_t M, int64_t N, int64_t K, int flags) {
  return precision == AGNN_GEMM_F16X3 && a_layout == AGNN_LAYOUT_K_MAJOR && M >= 4096 && N >= 256 && K >= 64 &&
         !(flags & (AGNN_GEMM_ACCUMULATE | AGNN_GEMM_OUT_BF16));
}

extern "C" int agnn_gemm_pair(int b_layout, int64_t M, int64_t N, int64_t K, const void* a_hi, const void* a_lo,
                              int64_t lda, const float* amax_a, const void* b_hi, const void* b_lo, int64_t ldb,
                              const float* amax_b, float* c, int64_t ldc, const float* bias, int flags, float* amax_out,
                              agnn_stream_t stream) {
  if (!agnn_gemm_pair_supported(AGNN_GEMM_F16X3, AGNN_LAYOUT_K_MAJOR, M, N, K, flags))
    return fail(AGNN_ERR_UNSUPPORTED, "gemm_pair: F16X3, K-major A, M >= 4096, N >= 256, no accumulate / bf16 output");
  if (!a_hi || !a_lo || !b_hi || !b_lo || !amax_a || !amax_b || !c || M >= (1ll << 31) || N >= (1ll << 31) ||
      K >= (1ll << 31))
    return fail(AGNN_ERR_ARG, "gemm_pair: null operand or bad sizes");
  if ((lda * 2) % 16 || (ldb * 2) % 16 || (ldc * 4) % 16 || !aligned16(a_hi) || !aligned16(a_lo) || !aligned16(b_hi) ||
      !aligned16(b_lo) || !aligned16(c))
    return fail(AGNN_ERR_UNSUPPORTED, "gemm_pair: operands and C must be 16-byte aligned with 16-byte multiple row strides");
  const bool b_mn = b_layout == AGNN_LAYOUT_MN_MAJOR;
  Gemm2Params p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.k_blocks = (int)ceil_div(K, kBlockK);
  p.tiles_m = (int)ceil_div(M, 256);
  p.tiles_n = (int)ceil_div(N, kBlockN);
  p.chain_blocks = 2;
  p.bias = bias; p.flags = flags; p.amax_a = amax_a; p.amax_b = amax_b; p.amax_out = amax_out;
  const void* a_parts[2] = {a_hi, a_lo};
  const void* b_parts[2] = {b_hi, b_lo};
  int rc;
  for (int i = 0; i < 2; ++i) {
    if ((rc = make_map2(&p.map_a[i], a_parts[i], true, M, K, lda, kBlockK, kBlockM))) return rc;
    rc = b_mn ? make_map2(&p.map_b[i], b_parts[i], true, K, N, ldb, kChunk, kBlockK)
              : make_map2(&p.map_b[i], b_parts[i], true, N, K, ldb, kBlockK, kHalfN);
    if (rc) return rc;
  }
  if ((rc = make_map2(&p.map_c, c, false, M, N, ldc, kStoreBox, kStoreBox))) return rc;
  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
  const int pairs = (int)(tiles < kNumSM / 2 ? tiles : kNumSM / 2);
  cudaStream_t st = (cudaStream_t)stream;
  return b_mn ? launch2<true>(p, 2 * pairs, st) : launch2<false>(p, 2 * pairs, st);
}
