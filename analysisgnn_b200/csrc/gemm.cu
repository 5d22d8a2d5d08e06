// Dense projections on the 5th-generation tensor cores: tcgen05.mma with TMEM accumulators,
// TMA-fed shared-memory operands, mbarrier pipelines, one persistent CTA per SM.
//
// Replaces the nn.Linear calls of the hot path -- SageConvScatter's neigh_linear / linear
// (analysisgnn/models/core/gnn.py:65, 75), PyG SAGEConv's lin_l / lin_r and HGT's kqv / out
// projections (third party), project_dict (analysisgnn/models/analysis.py:429-443) -- and the two
// GEMMs of their backward (grad-input, grad-weight).  See include/agnn.h (agnn_gemm).
//
// Precision modes
//   AGNN_GEMM_TF32X3  fp32 parity mode.  tcgen05 has no fp32-input MMA, so every fp32 operand is
//                     given as two TF32-exact matrices (hi = rna_tf32(x), lo = rna_tf32(x - hi),
//                     agnn_split_tf32) and D = Ahi*Bhi + Ahi*Blo + Alo*Bhi accumulates in fp32 TMEM.
//   AGNN_GEMM_TF32    one product (hi only): 1e-3 relative, for experiments.
//   AGNN_GEMM_BF16    bf16 operands, fp32 accumulation, bf16 or fp32 output: the stated bf16 mode.
//   AGNN_GEMM_F16X3   fp32 parity mode on the f16 MMA (twice the TF32 rate, half the operand bytes): every operand is
//                     hi = fp16(s x), lo = fp16(s x - hi) with one power-of-two scale s per tensor derived from its
//                     amax (agnn_amax, agnn_split_f16; max |s x| in [2^13, 2^14)).  fp16 has TF32's 11-bit significand,
//                     so the three products carry the same 22 bits as TF32X3 for elements down to 2^-17 of the
//                     tensor's amax (fixed point 2^-25 of the scaled range below); the epilogue undoes s_a s_b exactly.
//
// Operand layouts (both handled by the UMMA descriptors, no transposition pass):
//   K-major  : matrix stored [MN, K] row-major (activations as A, nn.Linear weights as B)
//   MN-major : matrix stored [K, MN] row-major (weights as B in grad-input; both operands in grad-weight)
//
// Grouped launches (agnn_gemm_grouped): up to AGNN_GEMM_MAX_GROUP independent problems of one precision / layout
// combination share ONE persistent launch -- the per-node-type projections (project_dict), the task heads, the
// destination types of a message-passing layer, the directions of a GRU layer; the CTAs draw the tiles of all problems,
// in order, from a per-launch device counter (GemmGroup::sched; by CTA index without one).  Split-K partials are combined INSIDE the launch: the CTA that stores the last partial of an
// output tile (a ticket counter per tile) adds the partials in split order -- deterministic, no second kernel.
// The epilogue can also emit max |C| (amax_out) so that a consumer that needs the fp16 operand scale of C does not
// re-read it.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected
// lane), warps 2-5 = epilogue (TMEM -> registers -> global, each warp its own 32 TMEM lanes).
// The accumulator is double buffered in TMEM (2 x 128 columns) so the epilogue of tile i overlaps
// the main loop of tile i+1.  Tile 128 x 128, K block = 128 bytes per row (32 tf32 / 64 bf16),
// 128-byte swizzle everywhere.
#include <cuda.h>

#include <atomic>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace agnn {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;
constexpr int kRowBytes = 128;              // one swizzle row = the K (or MN) extent of a tile row
constexpr int kTileBytes = kBlockM * kRowBytes;  // 16 KB per operand tile
constexpr int kThreads = 192;
constexpr int kTmemCols = 2 * kBlockN;
constexpr int kSmemBudget = 200 * 1024;      // operand stages
constexpr int kStoreBox = 32;                // epilogue staging: one warp's 32 rows x 32 fp32 columns ...
constexpr int kStoreBufBytes = kStoreBox * 128;          // ... = 4 KB, 128-byte swizzled, stored by TMA
constexpr int kStoreBytes = 4 * 2 * kStoreBufBytes;      // 4 epilogue warps x 2 buffers
constexpr int kSchedSlots = 4;               // work-item ids in flight between the producer and the other roles

struct GemmParams {
  CUtensorMap map_a[2];  // hi, lo
  CUtensorMap map_b[2];
  CUtensorMap map_c;     // fp32 C, box 32 x 32 (valid when tma_store)
  int tma_store;         // epilogue: registers -> swizzled smem -> TMA store / reduce-add
  int M, N, K;
  int k_blocks;          // K blocks in total
  int k_blocks_per_split;
  int split_k;
  int tiles_m, tiles_n;
  void* out;             // [split_k][tiles_m * 128][ld_part] when split_k > 1 (workspace), else C
  int64_t ldc;
  int64_t split_stride;  // elements between split partials
  void* c_final;         // C (the in-kernel split-K reduction writes it)
  int64_t ldc_final;
  const float* bias;
  int flags;
  const float* amax_a;   // F16X3: device scalars the operand scales derive from (null = unscaled operands)
  const float* amax_b;
  float* amax_out;       // optional: *amax_out = max(*amax_out, max |C|)
  int* tickets;          // split_k > 1: one counter per output tile (zero before the launch, zero after it), or null =
                         // the partials are reduced by splitk_reduce_kernel after the launch
};

constexpr int kMaxGroup = AGNN_GEMM_MAX_GROUP;

struct GemmGroup {
  int n_prob;
  int chain_blocks;      // K blocks accumulated in TMEM before the sum is promoted to fp32 registers
  int total_tiles;                 // GEMM work items (tile x split) of all problems
  int* sched;                      // work-item counter of this launch (zero before and after it), or null = items are
                                   // dealt round robin by CTA index
  int tile_start[kMaxGroup + 1];   // first work item (tile x split) of every problem
  // split-K with ticket counters: the CTA that stores the LAST partial of an output tile adds the tile's partials in
  // split order and finishes C.  Nobody ever waits for another CTA (two grids spinning on each other's unscheduled
  // CTAs from two streams could deadlock), the adds keep 32 independent 16-byte loads in flight per thread, and the
  // producer / MMA warps of that CTA go on with its next item meanwhile.
  GemmParams prob[kMaxGroup];
};

enum { kFmtTF32 = 0, kFmtBF16 = 1, kFmtF16 = 2 };

// ---------------------------------------------------------------- PTX wrappers
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// C tile out of shared memory: plain store, or element-wise fp32 add into global memory (one writer per element)
template <bool REDUCE>
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  if constexpr (REDUCE) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
  }
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

template <int FMT>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if constexpr (FMT != kFmtTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
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

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout SWIZZLE_128B (2) << 61
// layout: 2 = SWIZZLE_128B (16-byte swizzle atoms), 1 = SWIZZLE_128B_BASE32B (32-byte atoms; the only
// layout tcgen05 accepts for MN-major 32-bit operands -- TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
// (assembled in the MMA issue loop: constant high word, low word = address >> 4 | LBO >> 4 << 16)

// instruction descriptor (cute::UMMA::InstrDescriptor), D = fp32
// operand format codes (cute::UMMA::F16F32Format): F16 = 0, BF16 = 1, TF32 = 2
__host__ __device__ constexpr uint32_t instr_desc(int fmt, bool a_mn, bool b_mn, int m, int n) {
  return (1u << 4) | ((fmt == kFmtTF32 ? 2u : fmt == kFmtBF16 ? 1u : 0u) << 7) |
         ((fmt == kFmtTF32 ? 2u : fmt == kFmtBF16 ? 1u : 0u) << 10) | ((a_mn ? 1u : 0u) << 15) |
         ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// work item -> (problem, split, tile origin)
struct Item {
  int g, split, mn, m0, n0, kb0, kb1;
};
__device__ __forceinline__ Item decode_item(const GemmGroup& grp, int item) {
  Item it;
  int g = 0;
  while (g + 1 < grp.n_prob && item >= grp.tile_start[g + 1]) ++g;
  const GemmParams& p = grp.prob[g];
  const int local = item - grp.tile_start[g];
  const int per_split = p.tiles_m * p.tiles_n;
  it.g = g;
  it.split = local / per_split;
  it.mn = local - it.split * per_split;
  it.m0 = (it.mn / p.tiles_n) * kBlockM;
  it.n0 = (it.mn % p.tiles_n) * kBlockN;
  it.kb0 = it.split * p.k_blocks_per_split;
  it.kb1 = min(it.kb0 + p.k_blocks_per_split, p.k_blocks);
  return it;
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // the 4 epilogue warps

template <int FMT, bool A_MN, bool B_MN, int TERMS>
__global__ void __launch_bounds__(kThreads, 1) gemm_kernel(const __grid_constant__ GemmGroup grp) {
  constexpr bool BF16 = FMT != kFmtTF32;       // 2-byte operands (bf16 or fp16): same tiles and descriptors
  constexpr int kElem = BF16 ? 2 : 4;
  constexpr int kBlockK = kRowBytes / kElem;   // 32 tf32 / 64 bf16
  constexpr int kUmmaK = 32 / kElem;           // 8 tf32 / 16 bf16
  constexpr int kParts = TERMS == 3 ? 2 : 1;   // hi (+ lo)
  constexpr int kStageBytes = 2 * kParts * kTileBytes;
  constexpr int kStages = kSmemBudget / kStageBytes;
  constexpr int kChunk = kRowBytes / kElem;    // MN elements per 128-byte row of an MN-major tile
  constexpr int kChunks = kBlockM / kChunk;    // TMA boxes per MN-major tile
  constexpr uint32_t kIdesc = instr_desc(FMT, A_MN, B_MN, kBlockM, kBlockN);
  constexpr uint32_t kMnSbo = BF16 ? 1024 : 512;
  constexpr uint64_t kMnLayout = BF16 ? 2 : 1;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* store_base = smem + kStages * kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(store_base + kStoreBytes);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* sched_full = acc_empty + 2;
  uint64_t* sched_empty = sched_full + kSchedSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_empty + kSchedSlots);
  int* last_flag = reinterpret_cast<int*>(tmem_slot + 1);
  volatile int* sched_item = last_flag + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = grp.total_tiles;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < grp.n_prob; ++g) {
      for (int i = 0; i < kParts; ++i) {
        prefetch_tmap(&grp.prob[g].map_a[i]);
        prefetch_tmap(&grp.prob[g].map_b[i]);
      }
      if (grp.prob[g].tma_store) prefetch_tmap(&grp.prob[g].map_c);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);   // one arrive per epilogue warp
    }
    for (int s = 0; s < kSchedSlots; ++s) {
      mbar_init(&sched_full[s], 1);
      mbar_init(&sched_empty[s], 5);  // the MMA thread + one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // Work items are handed out by THIS thread: it draws the next item (a device-wide counter, or CTA index +
      // k * grid without one), publishes the id to the MMA and epilogue warps through a small ring, and loads it.
      // With the counter a CTA that starts late (the GRU kernels of the side stream hold whole SMs for a
      // millisecond, so part of a 148-CTA grid waits for a free SM) simply takes fewer items instead of working
      // through a fixed 1/148 share after everybody else has finished.  -1 ends the roles.
      int sslot = 0;
      uint32_t sphase = 0;
      int next_static = blockIdx.x;
      auto draw = [&]() -> int {
        if (grp.sched) {
          const int got = atomicAdd(grp.sched, 1);
          if (got < n_items) return got;
          // every CTA draws exactly one id >= n_items; the last of them leaves the counter at zero for the next launch
          if (got == n_items + (int)gridDim.x - 1) atomicExch(grp.sched, 0);
          return -1;
        }
        const int got = next_static < n_items ? next_static : -1;
        next_static += gridDim.x;
        return got;
      };
      int item = draw();
      while (true) {
        mbar_wait(&sched_empty[sslot], sphase ^ 1);
        sched_item[sslot] = item;
        mbar_arrive(&sched_full[sslot]);
        if (++sslot == kSchedSlots) { sslot = 0; sphase ^= 1; }
        if (item < 0) break;
        const Item it = decode_item(grp, item);
        item = draw();                               // in flight while this item's loads are issued
        const GemmParams& p = grp.prob[it.g];
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * kStageBytes;
          mbar_expect_tx(&full[stage], kStageBytes);
          const int k0 = kb * kBlockK;
#pragma unroll
          for (int part = 0; part < kParts; ++part) {
            uint8_t* a_dst = st + part * kTileBytes;
            uint8_t* b_dst = st + (kParts + part) * kTileBytes;
            if constexpr (A_MN) {
#pragma unroll
              for (int c = 0; c < kChunks; ++c)
                tma_load_2d(a_dst + c * (kBlockK * kRowBytes), &p.map_a[part], &full[stage], it.m0 + c * kChunk, k0);
            } else {
              tma_load_2d(a_dst, &p.map_a[part], &full[stage], k0, it.m0);
            }
            if constexpr (B_MN) {
#pragma unroll
              for (int c = 0; c < kChunks; ++c)
                tma_load_2d(b_dst + c * (kBlockK * kRowBytes), &p.map_b[part], &full[stage], it.n0 + c * kChunk, k0);
            } else {
              tma_load_2d(b_dst, &p.map_b[part], &full[stage], k0, it.n0);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The tensor core adds into the fp32 TMEM accumulator with truncation (measured: -0.5 ulp per MMA,
    // i.e. a bias that grows linearly with the chain).  A chain is therefore limited to
    // grp.chain_blocks K blocks; the epilogue warps sum the chains in registers with round-to-nearest.
    // One elected thread runs the whole issue loop (waits included): the loop is ~40 instructions per K block,
    // against ~130 when every lane walked it and the descriptors were rebuilt in 64-bit arithmetic per MMA --
    // at 12 MMAs per K block the issue path, not the tensor pipe, was setting the pace.
    if (elect_one()) {
      // descriptor high words are constants; low word = (address >> 4) | (LBO >> 4) << 16
      constexpr uint32_t kHiK = (1024u >> 4) | (1u << 14) | (2u << 29);                     // K-major, SWIZZLE_128B
      constexpr uint32_t kHiMn = (kMnSbo >> 4) | (1u << 14) | ((uint32_t)kMnLayout << 29);  // MN-major
      constexpr uint32_t kLboK = (16u >> 4) << 16;
      constexpr uint32_t kLboMn = (((uint32_t)(kBlockK * kRowBytes) >> 4) & 0x3FFF) << 16;
      constexpr uint32_t kStepA = A_MN ? (kUmmaK * kRowBytes) >> 4 : 32 >> 4;               // per UMMA K step
      constexpr uint32_t kStepB = B_MN ? (kUmmaK * kRowBytes) >> 4 : 32 >> 4;
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int cc = 0;                                    // chains issued so far (TMEM buffer = cc & 1)
      int sslot = 0;
      uint32_t sphase = 0;
      while (true) {
        mbar_wait(&sched_full[sslot], sphase);
        const int item = sched_item[sslot];
        mbar_arrive(&sched_empty[sslot]);
        if (++sslot == kSchedSlots) { sslot = 0; sphase ^= 1; }
        if (item < 0) break;
        const Item it = decode_item(grp, item);
        for (int c0 = it.kb0; c0 < it.kb1; c0 += grp.chain_blocks, ++cc) {
          const int c1 = min(c0 + grp.chain_blocks, it.kb1);
          const int buf = cc & 1;
          mbar_wait(&acc_empty[buf], ((cc >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * kBlockN;
          uint32_t accumulate = 0;
          for (int kb = c0; kb < c1; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t st = (smem_base + stage * kStageBytes) >> 4;
#pragma unroll
            for (int term = 0; term < TERMS; ++term) {
              // term 0: hi*hi, 1: hi*lo, 2: lo*hi
              const uint32_t a_lo = (st + (term == 2 ? 1 : 0) * (kTileBytes >> 4)) | (A_MN ? kLboMn : kLboK);
              const uint32_t b_lo = (st + (kParts + (term == 1 ? 1 : 0)) * (kTileBytes >> 4)) | (B_MN ? kLboMn : kLboK);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                // K-major: advance 32 bytes inside the 128-byte swizzle row; 8-row groups are 1024 B apart.
                // MN-major: one K step = kUmmaK rows of 128 bytes; MN chunks are kBlockK rows apart;
                // 32-bit operands use the 32-byte-atom swizzle (groups of 4 K rows, 512 B).
                const uint64_t da = ((uint64_t)(A_MN ? kHiMn : kHiK) << 32) | (a_lo + k * kStepA);
                const uint64_t db = ((uint64_t)(B_MN ? kHiMn : kHiK) << 32) | (b_lo + k * kStepB);
                umma<FMT>(tmem_d, da, db, kIdesc, accumulate);
                accumulate = 1;
              }
            }
            umma_commit(&empty[stage]);              // frees the smem stage when these MMAs retire
            if (kb == c1 - 1) umma_commit(&acc_full[buf]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;                  // TMEM lanes [32*quarter, +32) belong to this warp
    const int et = threadIdx.x - 64;               // 0..127 among the epilogue threads
    uint8_t* sbuf = store_base + quarter * 2 * kStoreBufBytes;
    int cc = 0;
    int sc = 0;                                    // staged chunks so far (buffer = sc & 1)
    int sslot = 0;
    uint32_t sphase = 0;
    while (true) {
      mbar_wait(&sched_full[sslot], sphase);
      const int item = sched_item[sslot];
      __syncwarp();                                // every lane has read the id before the slot is handed back
      if (lane == 0) mbar_arrive(&sched_empty[sslot]);
      if (++sslot == kSchedSlots) { sslot = 0; sphase ^= 1; }
      if (item < 0) break;
      const Item it = decode_item(grp, item);
      const GemmParams& p = grp.prob[it.g];
      const int m0 = it.m0, n0 = it.n0;
      const bool direct = p.split_k == 1;
      // F16X3: the operands were scaled by powers of two; dividing by them (exactly) comes before the bias
      const bool scaled = p.amax_a != nullptr;
      const float inv_a = scaled ? 1.f / f16_scale_of(__ldg(p.amax_a)) : 1.f;
      const float inv_b = scaled ? 1.f / f16_scale_of(__ldg(p.amax_b)) : 1.f;
      const int row = m0 + quarter * 32 + lane;
      // The bias is the first addend: its loads are in flight while the first chain is computed.
      float acc[kBlockN];
      if (direct && p.bias && !scaled) {
#pragma unroll
        for (int j = 0; j < kBlockN; ++j) acc[j] = n0 + j < p.N ? __ldg(p.bias + n0 + j) : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < kBlockN; ++j) acc[j] = 0.f;
      }
      for (int c0 = it.kb0; c0 < it.kb1; c0 += grp.chain_blocks, ++cc) {
        const int buf = cc & 1;
        mbar_wait(&acc_full[buf], (cc >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + buf * kBlockN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
        for (int c = 0; c < kBlockN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
      if (scaled) {
        if (direct && p.bias) {
#pragma unroll
          for (int j = 0; j < kBlockN; ++j)
            acc[j] = (acc[j] * inv_a) * inv_b + (n0 + j < p.N ? __ldg(p.bias + n0 + j) : 0.f);
        } else {
#pragma unroll
          for (int j = 0; j < kBlockN; ++j) acc[j] = (acc[j] * inv_a) * inv_b;
        }
      }
      if (direct && (p.flags & AGNN_GEMM_RELU) && !(p.flags & AGNN_GEMM_ACCUMULATE)) {
#pragma unroll
        for (int j = 0; j < kBlockN; ++j) acc[j] = fmaxf(acc[j], 0.f);
      }
      if (direct && p.amax_out && !(p.flags & AGNN_GEMM_ACCUMULATE)) {
        // max |C| over the valid part of this thread's row (rows >= M hold the bias only: not part of C)
        uint32_t mx = 0;
        if (row < p.M) {
#pragma unroll
          for (int j = 0; j < kBlockN; ++j)
            if (n0 + j < p.N) mx = max(mx, __float_as_uint(acc[j]) & 0x7fffffffu);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        if (lane == 0 && mx) atomicMax(reinterpret_cast<unsigned int*>(p.amax_out), mx);
      }
      if (p.tma_store) {
        // registers -> 128-byte-swizzled shared memory (this warp's 32 rows x 32 columns) -> one TMA store
        // (or fp32 reduce-add for AGNN_GEMM_ACCUMULATE); the map clips rows >= M and columns >= N.
#pragma unroll
        for (int c = 0; c < kBlockN / 32; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;
          float* v = acc + c * 32;
          uint8_t* buf = sbuf + (sc & 1) * kStoreBufBytes;
          ++sc;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          const uint32_t row_addr = smem_u32(buf) + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((j ^ (lane & 7)) << 4)),
                         "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                         : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (p.flags & AGNN_GEMM_ACCUMULATE) tma_store_2d<true>(&p.map_c, buf, col0, m0 + quarter * 32);
            else tma_store_2d<false>(&p.map_c, buf, col0, m0 + quarter * 32);
          }
        }
      } else if (direct) {
        if (row < p.M) {
#pragma unroll
          for (int c = 0; c < kBlockN / 32; ++c) {
            const int col0 = n0 + c * 32;
            if (col0 >= p.N) break;
            float* v = acc + c * 32;
            const int64_t off = (int64_t)row * p.ldc + col0;
            if (!(p.flags & AGNN_GEMM_OUT_BF16)) {
              float* o = static_cast<float*>(p.out) + off;
              if (p.flags & AGNN_GEMM_ACCUMULATE) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) v[j] += o[j];
                if (p.flags & AGNN_GEMM_RELU) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
              }
              if (col0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) o[j] = v[j];
              }
            } else {
              __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + off;
              if (p.flags & AGNN_GEMM_ACCUMULATE) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) v[j] += __bfloat162float(o[j]);
                if (p.flags & AGNN_GEMM_RELU) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
              }
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = __float2bfloat16_rn(v[j]);
            }
          }
        }
      } else {
        // ---- split-K: this item's partial tile -> workspace [split][tiles_m * 128][ld_part] (rows padded to whole
        // tiles, so every row of the tile has a slot).  Staged through this warp's shared-memory box so that 8 lanes
        // write 128 contiguous bytes.
        float* part = static_cast<float*>(p.out) + (int64_t)it.split * p.split_stride;
        uint8_t* buf = sbuf;
#pragma unroll
        for (int c = 0; c < kBlockN / 32; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;
          float* v = acc + c * 32;
          if (sc) {                                 // a TMA store of an earlier direct tile may still read the box
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            sc = 0;
          }
          __syncwarp();
          const uint32_t row_addr = smem_u32(buf) + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((j ^ (lane & 7)) << 4)),
                         "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                         : "memory");
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3), q = lane & 7;
            float4 t;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                         : "r"(smem_u32(buf) + r * 128 + ((q ^ (r & 7)) << 4)));
            if (m0 + quarter * 32 + r < p.M && col0 + q * 4 < p.N)
              *reinterpret_cast<float4*>(part + (int64_t)(m0 + quarter * 32 + r) * p.ldc + col0 + q * 4) = t;
          }
          __syncwarp();
        }
        if (p.tickets) {
          __threadfence();                          // this CTA's partial is visible before its ticket
          epi_barrier();
          if (et == 0) {
            const int old = atomicAdd(p.tickets + it.mn, 1);
            *last_flag = (old == p.split_k - 1) ? 1 : 0;
          }
          epi_barrier();
          const bool last = *last_flag != 0;
          epi_barrier();                            // everyone has read the flag before the next item rewrites it
          if (last) {
            __threadfence();
            const int qv = et & 31;                 // float4 column of the tile
            const int gcol = n0 + qv * 4;
            uint32_t mx = 0;
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias && gcol < p.N) {
              bias4.x = __ldg(p.bias + gcol);
              if (gcol + 1 < p.N) bias4.y = __ldg(p.bias + gcol + 1);
              if (gcol + 2 < p.N) bias4.z = __ldg(p.bias + gcol + 2);
              if (gcol + 3 < p.N) bias4.w = __ldg(p.bias + gcol + 3);
            }
            const int rows_here = min(kBlockM, p.M - m0);
            // thread (w, qv) owns rows w, w + 4, ...; four of its rows x eight splits = 32 loads in flight
#pragma unroll 1
            for (int r0 = et >> 5; r0 < rows_here && gcol < p.N; r0 += 16) {
              float4 sum[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) sum[j] = bias4;
              const float* src = static_cast<const float*>(p.out) + (int64_t)(m0 + r0) * p.ldc + gcol;
#pragma unroll 1
              for (int sp = 0; sp < p.split_k; sp += 8) {
                float4 t[4][8];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                  for (int u = 0; u < 8; ++u) {
                    const bool on = r0 + 4 * j < rows_here && sp + u < p.split_k;
                    t[j][u] = on ? __ldcg(reinterpret_cast<const float4*>(src + (int64_t)(4 * j) * p.ldc +
                                                                           (int64_t)(sp + u) * p.split_stride))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                  }
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                  for (int u = 0; u < 8; ++u) {       // split order: deterministic
                    sum[j].x += t[j][u].x; sum[j].y += t[j][u].y; sum[j].z += t[j][u].z; sum[j].w += t[j][u].w;
                  }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int grow = m0 + r0 + 4 * j;
                if (r0 + 4 * j >= rows_here) continue;
                float v[4] = {sum[j].x, sum[j].y, sum[j].z, sum[j].w};
                if (p.flags & AGNN_GEMM_OUT_BF16) {
                  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.c_final) + (int64_t)grow * p.ldc_final + gcol;
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (gcol + e < p.N) {
                      float x = v[e];
                      if (p.flags & AGNN_GEMM_ACCUMULATE) x += __bfloat162float(o[e]);
                      if (p.flags & AGNN_GEMM_RELU) x = fmaxf(x, 0.f);
                      mx = max(mx, __float_as_uint(x) & 0x7fffffffu);
                      o[e] = __float2bfloat16_rn(x);
                    }
                } else {
                  float* o = static_cast<float*>(p.c_final) + (int64_t)grow * p.ldc_final + gcol;
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (gcol + e < p.N) {
                      if (p.flags & AGNN_GEMM_ACCUMULATE) v[e] += o[e];
                      if (p.flags & AGNN_GEMM_RELU) v[e] = fmaxf(v[e], 0.f);
                      mx = max(mx, __float_as_uint(v[e]) & 0x7fffffffu);
                    }
                  if (gcol + 4 <= p.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                  } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                      if (gcol + e < p.N) o[e] = v[e];
                  }
                }
              }
            }
            if (p.amax_out) {
#pragma unroll
              for (int d = 16; d >= 1; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
              if (lane == 0 && mx) atomicMax(reinterpret_cast<unsigned int*>(p.amax_out), mx);
            }
            if (et == 0) p.tickets[it.mn] = 0;      // self-cleaning: the counters are zero again after the launch
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ---------------------------------------------------------------- split / reduce helpers
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, int64_t rows, int cols4,
                                                          int64_t ld_x, float* __restrict__ hi, float* __restrict__ lo,
                                                          int64_t ld_o) {
  const int64_t total = rows * cols4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld_x + c));
    const float in[4] = {v.x, v.y, v.z, v.w};
    float h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(in[k]));
      h[k] = __uint_as_float(hb);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(in[k] - h[k]));
      l[k] = __uint_as_float(lb);
    }
    *reinterpret_cast<float4*>(hi + r * ld_o + c) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo + r * ld_o + c) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// C (+)= bias + sum_s partial[s] for every split-K problem of a grouped launch, in split order (deterministic),
// optional ReLU / bf16 output / amax.  One launch for the whole group: blocks are dealt to the problems in proportion
// to their output size; a thread owns one float4 of C and keeps 8 partial loads in flight.
struct ReduceGroup {
  int n_prob;
  int blk_start[kMaxGroup + 1];
  struct P {
    const float* part; int split_k; int64_t split_stride, ld_p;
    void* out; int64_t ldc; const float* bias; int flags; int M, N; float* amax_out;
  } prob[kMaxGroup];
};

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const __grid_constant__ ReduceGroup grp) {
  int g = 0;
  while (g + 1 < grp.n_prob && (int)blockIdx.x >= grp.blk_start[g + 1]) ++g;
  const ReduceGroup::P& p = grp.prob[g];
  const int n4 = (p.N + 3) / 4;
  const int64_t total = (int64_t)p.M * n4;
  const int64_t nblk = grp.blk_start[g + 1] - grp.blk_start[g];
  uint32_t mx = 0;
  for (int64_t i = (int64_t)(blockIdx.x - grp.blk_start[g]) * 256 + threadIdx.x; i < total; i += nblk * 256) {
    const int64_t m = i / n4;
    const int n = (int)(i - m * n4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) {
      acc.x = __ldg(p.bias + n);
      if (n + 1 < p.N) acc.y = __ldg(p.bias + n + 1);
      if (n + 2 < p.N) acc.z = __ldg(p.bias + n + 2);
      if (n + 3 < p.N) acc.w = __ldg(p.bias + n + 3);
    }
    const float* src = p.part + m * p.ld_p + n;
    int sp = 0;
    for (; sp + 8 <= p.split_k; sp += 8) {
      float4 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(sp + u) * p.split_stride));
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w; }
    }
    for (; sp < p.split_k; ++sp) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src + (int64_t)sp * p.split_stride));
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
    if (p.flags & AGNN_GEMM_OUT_BF16) {
      __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + m * p.ldc + n;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (n + e < p.N) {
          float x = v[e];
          if (p.flags & AGNN_GEMM_ACCUMULATE) x += __bfloat162float(o[e]);
          if (p.flags & AGNN_GEMM_RELU) x = fmaxf(x, 0.f);
          mx = max(mx, __float_as_uint(x) & 0x7fffffffu);
          o[e] = __float2bfloat16_rn(x);
        }
    } else {
      float* o = static_cast<float*>(p.out) + m * p.ldc + n;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (n + e < p.N) {
          if (p.flags & AGNN_GEMM_ACCUMULATE) v[e] += o[e];
          if (p.flags & AGNN_GEMM_RELU) v[e] = fmaxf(v[e], 0.f);
          mx = max(mx, __float_as_uint(v[e]) & 0x7fffffffu);
        }
      if (n + 4 <= p.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < p.N) o[e] = v[e];
      }
    }
  }
  if (p.amax_out) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(reinterpret_cast<unsigned int*>(p.amax_out), mx);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// 2-D map over a row-major matrix [rows, cols] (cols contiguous), box = box_cols x box_rows, 128B swizzle
int make_map(CUtensorMap* map, const void* ptr, int fmt, int64_t rows, int64_t cols, int64_t ld, int box_cols,
             int box_rows, bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(AGNN_ERR_CUDA, "gemm: cuTensorMapEncodeTiled is not available from this driver");
  const bool bf16 = fmt != kFmtTF32;   // 2-byte elements
  const int eb = bf16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * eb};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(map, fmt == kFmtBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : fmt == kFmtF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                                          : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   (mn_major && !bf16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(AGNN_ERR_CUDA, "gemm: cuTensorMapEncodeTiled failed (%d)", (int)rc);
  return AGNN_OK;
}

// Work-item counters of the launches (GemmGroup::sched).  Every launch leaves its counter at zero, so a counter can be
// reused as soon as the launch is over; launches that may run at the same time (two streams, or two branches of a
// captured graph) must not share one, hence a ring handed out in call order: a clash needs two launches half a ring of
// calls apart to overlap.  Launches recorded into a CUDA graph keep their counter for the life of the graph, so they
// draw from their own half of the ring: an eager launch on another stream can never meet a replaying graph's counter.
// A static device array: no allocation, nothing to free, capturable.
constexpr int kSchedCounters = 16384;
__device__ int g_sched_counters[kSchedCounters];

int* next_sched_counter(cudaStream_t st) {
  static int* base = [] {
    void* p = nullptr;
    return cudaGetSymbolAddress(&p, g_sched_counters) == cudaSuccess ? static_cast<int*>(p) : nullptr;
  }();
  static std::atomic<unsigned> next_eager{0}, next_captured{0};
  if (!base) return nullptr;
  constexpr unsigned kHalf = kSchedCounters / 2;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  const bool captured = cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive;
  return captured ? base + kHalf + next_captured.fetch_add(1, std::memory_order_relaxed) % kHalf
                  : base + next_eager.fetch_add(1, std::memory_order_relaxed) % kHalf;
}

bool sched_dynamic() {      // AGNN_GEMM_SCHED=static: items dealt round robin by CTA index (the round-1 behaviour)
  static const bool on = [] { const char* e = getenv("AGNN_GEMM_SCHED"); return !(e && !strcmp(e, "static")); }();
  return on;
}

template <int FMT, bool A_MN, bool B_MN, int TERMS>
int launch(const GemmGroup& grp, int grid, cudaStream_t st) {
  constexpr int kParts = TERMS == 3 ? 2 : 1;
  constexpr int kStageBytes = 2 * kParts * kTileBytes;
  constexpr int kStages = kSmemBudget / kStageBytes;
  constexpr int smem = kStages * kStageBytes + kStoreBytes + 1024 /*align*/ + 256 /*barriers*/;
  auto kern = gemm_kernel<FMT, A_MN, B_MN, TERMS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return check_launch("gemm: cudaFuncSetAttribute");
    configured = true;
  }
  kern<<<grid, kThreads, smem, st>>>(grp);
  return check_launch("gemm");
}

template <int FMT, int TERMS>
int dispatch_layout(bool a_mn, bool b_mn, const GemmGroup& grp, int grid, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch<FMT, false, false, TERMS>(grp, grid, st);
  if (!a_mn && b_mn) return launch<FMT, false, true, TERMS>(grp, grid, st);
  if (a_mn && b_mn) return launch<FMT, true, true, TERMS>(grp, grid, st);
  return launch<FMT, true, false, TERMS>(grp, grid, st);
}

inline int64_t part_ld(int64_t N) { return ceil_div(N, kBlockN) * kBlockN; }
inline int64_t part_rows(int64_t M) { return ceil_div(M, kBlockM) * kBlockM; }

// fills grp.prob[i] from a problem description (validation, tensor maps, split bookkeeping)
int setup_problem(GemmParams& p, int precision, bool a_mn, bool b_mn, const agnn_gemm_problem_t& q, int* tickets) {
  const int64_t M = q.M, N = q.N, K = q.K;
  const bool bf16 = precision == AGNN_GEMM_BF16;
  const bool three = precision == AGNN_GEMM_TF32X3 || precision == AGNN_GEMM_F16X3;
  const int fmt = bf16 ? kFmtBF16 : precision == AGNN_GEMM_F16X3 ? kFmtF16 : kFmtTF32;
  const int eb = fmt == kFmtTF32 ? 4 : 2, block_k = kRowBytes / eb, chunk = kRowBytes / eb;
  if ((q.amax_a || q.amax_b) && (precision != AGNN_GEMM_F16X3 || !q.amax_a || !q.amax_b))
    return fail(AGNN_ERR_ARG, "gemm: operand scales belong to the F16X3 mode and come in pairs");
  if (M <= 0 || N <= 0 || K <= 0 || !q.c || M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31))
    return fail(AGNN_ERR_ARG, "gemm: bad sizes");
  if (!q.a_hi || !q.b_hi || (three && (!q.a_lo || !q.b_lo)))
    return fail(AGNN_ERR_ARG, "gemm: null operand (the three-product modes need the hi and lo parts of both operands)");
  if ((q.lda * eb) % 16 || (q.ldb * eb) % 16 || !aligned16(q.a_hi) || !aligned16(q.b_hi) ||
      (q.a_lo && !aligned16(q.a_lo)) || (q.b_lo && !aligned16(q.b_lo)))
    return fail(AGNN_ERR_UNSUPPORTED, "gemm: operands must be 16-byte aligned with 16-byte multiple row strides");
  if ((q.flags & AGNN_GEMM_OUT_BF16) && !bf16) return fail(AGNN_ERR_ARG, "gemm: bf16 output needs the bf16 mode");
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.k_blocks = (int)ceil_div(K, block_k);
  int split_k = q.split_k < 1 ? 1 : q.split_k;
  if (split_k > p.k_blocks) split_k = p.k_blocks;
  p.k_blocks_per_split = (int)ceil_div(p.k_blocks, split_k);
  split_k = (int)ceil_div(p.k_blocks, p.k_blocks_per_split);
  p.split_k = split_k;
  if (q.amax_out && (q.flags & AGNN_GEMM_ACCUMULATE) && split_k == 1)
    return fail(AGNN_ERR_ARG, "gemm: amax_out is not available together with AGNN_GEMM_ACCUMULATE");
  p.amax_a = q.amax_a;
  p.amax_b = q.amax_b;
  p.amax_out = q.amax_out;
  p.tiles_m = (int)ceil_div(M, kBlockM);
  p.tiles_n = (int)ceil_div(N, kBlockN);
  p.bias = q.bias;
  p.flags = q.flags;
  p.c_final = q.c;
  p.ldc_final = q.ldc;
  if (split_k > 1) {
    const int64_t ld_part = part_ld(N);
    const size_t need = (size_t)split_k * (size_t)part_rows(M) * (size_t)ld_part * sizeof(float);
    if (!q.workspace || q.workspace_bytes < need || !aligned16(q.workspace))
      return fail(AGNN_ERR_WORKSPACE, "gemm: split-K workspace %zu < %zu bytes", q.workspace_bytes, need);
    p.out = q.workspace; p.ldc = ld_part; p.split_stride = part_rows(M) * ld_part;
    p.tickets = tickets;
  } else {
    p.out = q.c; p.ldc = q.ldc; p.split_stride = 0;
  }
  int rc;
  const void* a_parts[2] = {q.a_hi, q.a_lo};
  const void* b_parts[2] = {q.b_hi, q.b_lo};
  const int parts = three ? 2 : 1;
  for (int i = 0; i < parts; ++i) {
    // K-major: [MN, K] row-major, box = block_k x 128 rows.  MN-major: [K, MN] row-major, box = chunk x block_k rows.
    rc = a_mn ? make_map(&p.map_a[i], a_parts[i], fmt, K, M, q.lda, chunk, block_k, true)
              : make_map(&p.map_a[i], a_parts[i], fmt, M, K, q.lda, block_k, kBlockM, false);
    if (rc) return rc;
    rc = b_mn ? make_map(&p.map_b[i], b_parts[i], fmt, K, N, q.ldb, chunk, block_k, true)
              : make_map(&p.map_b[i], b_parts[i], fmt, N, K, q.ldb, block_k, kBlockN, false);
    if (rc) return rc;
  }
  const bool acc_relu = (q.flags & AGNN_GEMM_ACCUMULATE) && (q.flags & AGNN_GEMM_RELU);
  if (split_k == 1 && !(q.flags & AGNN_GEMM_OUT_BF16) && !acc_relu && (q.ldc * 4) % 16 == 0 && aligned16(q.c)) {
    rc = make_map(&p.map_c, q.c, kFmtTF32, M, N, q.ldc, kStoreBox, kStoreBox, false);
    if (rc) return rc;
    p.tma_store = 1;
  }
  return AGNN_OK;
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_split_tf32(const float* x, int64_t rows, int64_t cols, int64_t ld_x, float* hi, float* lo,
                               int64_t ld_out, agnn_stream_t stream) {
  if (rows < 0 || cols < 0 || (cols % 4) || (ld_x % 4) || (ld_out % 4) || !aligned16(x) || !aligned16(hi) || !aligned16(lo))
    return fail(AGNN_ERR_ARG, "split_tf32: needs 16-byte aligned rows and a column count multiple of 4");
  if (rows == 0 || cols == 0) return AGNN_OK;
  int64_t blocks = ceil_div(rows * (cols / 4), 256);
  if (blocks > kNumSM * 16) blocks = kNumSM * 16;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, (int)(cols / 4), ld_x, hi, lo, ld_out);
  return check_launch("split_tf32");
}

// *amax = max(*amax, max |x|): non-negative floats order like their bit patterns, so one integer atomicMax per block
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, int64_t rows, int cols4, int64_t ld_x,
                                                    float* __restrict__ amax) {
  __shared__ uint32_t red[8];
  const int64_t total = rows * cols4;
  uint32_t m = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld_x + c));
    m = max(max(m, __float_as_uint(v.x) & 0x7fffffffu), __float_as_uint(v.y) & 0x7fffffffu);
    m = max(max(m, __float_as_uint(v.z) & 0x7fffffffu), __float_as_uint(v.w) & 0x7fffffffu);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = max(m, red[w]);
    if (m) atomicMax(reinterpret_cast<unsigned int*>(amax), m);
  }
}

__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, int64_t rows, int cols4,
                                                         int64_t ld_x, const float* __restrict__ amax,
                                                         __half* __restrict__ hi, __half* __restrict__ lo, int64_t ld_o,
                                                         float drop_p, const uint64_t* __restrict__ rng,
                                                         uint32_t rng_stream, int seq_len, int shift) {
  const float s = f16_scale_of(__ldg(amax));
  const int64_t total = rows * cols4;
  const bool drop = drop_p > 0.f && rng;
  const float keep_scale = drop ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t thr = dropout_threshold(drop_p);
  const uint64_t key = drop ? dropout_key(rng, rng_stream) : 0ull;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    float in[4] = {0.f, 0.f, 0.f, 0.f};
    // shift != 0: output row r holds input row r - shift of the SAME sequence (rows come in sequences of seq_len),
    // zeros where that row does not exist -- the h_{t-1} (or h_{t+1}) operand of a GRU's recurrent weight gradient
    const int64_t rs = r - shift;
    if (shift == 0 || (rs >= 0 && rs < rows && rs / seq_len == r / seq_len)) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + rs * ld_x + c));
      in[0] = v.x; in[1] = v.y; in[2] = v.z; in[3] = v.w;
    }
    if (drop) {                                   // same mask as agnn_dropout_apply on [rows, cols]
      const uint64_t bits = dropout_bits(key, (uint64_t)i);
#pragma unroll
      for (int e = 0; e < 4; ++e) in[e] = dropout_keep(bits, e, thr) ? in[e] * keep_scale : 0.f;
    }
    uint2 h, l;
    f16_pair4(in, s, h, l);
    *reinterpret_cast<uint2*>(hi + r * ld_o + c) = h;
    *reinterpret_cast<uint2*>(lo + r * ld_o + c) = l;
  }
}

// ---- amax + fp16 pair of SEVERAL small matrices (the weights of a grouped launch) in one launch ----------------------
// One cluster of 8 CTAs per matrix: every CTA takes the maximum over its share, the eight partial maxima meet through
// distributed shared memory (one cluster barrier), every CTA then splits its share with the common scale.  Replaces
// two launches per weight and step (agnn_amax + agnn_split_f16) by one launch per GEMM group.
constexpr int kSplitCluster = 8;

struct SplitMulti {
  int n;
  struct T {
    const float* x; int64_t rows; int cols4; int64_t ld_x;
    __half* hi; __half* lo; int64_t ld_o; float* amax;
  } t[AGNN_SPLIT_MULTI_MAX];
};

constexpr int kSplitThreads = 1024;

__global__ void __cluster_dims__(kSplitCluster, 1, 1) __launch_bounds__(kSplitThreads)
split_f16_multi_kernel(const __grid_constant__ SplitMulti p) {
  __shared__ uint32_t red[kSplitThreads / 32];
  __shared__ uint32_t cta_max;
  const SplitMulti::T& t = p.t[blockIdx.x / kSplitCluster];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int64_t total = t.rows * t.cols4;
  uint32_t m = 0;
#pragma unroll 4
  for (int64_t i = (int64_t)rank * kSplitThreads + threadIdx.x; i < total; i += (int64_t)kSplitCluster * kSplitThreads) {
    const int64_t r = i / t.cols4;
    const int c = (int)(i - r * t.cols4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(t.x + r * t.ld_x + c));
    m = max(max(m, __float_as_uint(v.x) & 0x7fffffffu), __float_as_uint(v.y) & 0x7fffffffu);
    m = max(max(m, __float_as_uint(v.z) & 0x7fffffffu), __float_as_uint(v.w) & 0x7fffffffu);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kSplitThreads / 32; ++w) m = max(m, red[w]);
    cta_max = m;
  }
  // all eight partial maxima are written ...
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  uint32_t all = 0;
#pragma unroll
  for (uint32_t r = 0; r < kSplitCluster; ++r) {
    uint32_t remote, v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(&cta_max)), "r"(r));
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
    all = max(all, v);
  }
  // ... and read: nobody leaves (or overwrites) before the others are done with its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const float amax = __uint_as_float(all);
  if (rank == 0 && threadIdx.x == 0) *t.amax = amax;
  const float s = f16_scale_of(amax);
#pragma unroll 4
  for (int64_t i = (int64_t)rank * kSplitThreads + threadIdx.x; i < total; i += (int64_t)kSplitCluster * kSplitThreads) {
    const int64_t r = i / t.cols4;
    const int c = (int)(i - r * t.cols4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(t.x + r * t.ld_x + c));
    const float in[4] = {v.x, v.y, v.z, v.w};
    uint2 h, l;
    f16_pair4(in, s, h, l);
    *reinterpret_cast<uint2*>(t.hi + r * t.ld_o + c) = h;
    *reinterpret_cast<uint2*>(t.lo + r * t.ld_o + c) = l;
  }
}

extern "C" int agnn_split_f16_multi(int n, const agnn_split_item_t* items, agnn_stream_t stream) {
  if (n < 0 || n > AGNN_SPLIT_MULTI_MAX || (n && !items))
    return fail(AGNN_ERR_ARG, "split_f16_multi: 0..%d matrices per launch, got %d", AGNN_SPLIT_MULTI_MAX, n);
  if (n == 0) return AGNN_OK;
  SplitMulti p;
  p.n = n;
  for (int i = 0; i < n; ++i) {
    const agnn_split_item_t& q = items[i];
    if (q.rows <= 0 || q.cols <= 0 || q.cols % 4 || (q.ld_x * 4) % 16 || (q.ld_out * 2) % 16 || !q.x || !q.hi || !q.lo ||
        !q.amax || !aligned16(q.x) || !aligned16(q.hi) || !aligned16(q.lo))
      return fail(AGNN_ERR_ARG, "split_f16_multi: matrix %d needs 16-byte aligned rows, a column count multiple of 4 and "
                                "non-null pointers", i);
    p.t[i].x = q.x; p.t[i].rows = q.rows; p.t[i].cols4 = (int)(q.cols / 4); p.t[i].ld_x = q.ld_x;
    p.t[i].hi = static_cast<__half*>(q.hi); p.t[i].lo = static_cast<__half*>(q.lo); p.t[i].ld_o = q.ld_out;
    p.t[i].amax = q.amax;
  }
  split_f16_multi_kernel<<<n * kSplitCluster, kSplitThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("split_f16_multi");
}

extern "C" int agnn_amax(const float* x, int64_t rows, int64_t cols, int64_t ld_x, float* amax, agnn_stream_t stream) {
  if (rows < 0 || cols < 0 || !amax) return fail(AGNN_ERR_ARG, "amax: bad arguments");
  if (rows == 0 || cols == 0) return AGNN_OK;
  if (cols % 4 || (ld_x * 4) % 16 || !aligned16(x))
    return fail(AGNN_ERR_ARG, "amax: needs 16-byte aligned rows and a column count multiple of 4");
  int64_t blocks = ceil_div(rows * (cols / 4), 256 * 4);
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  amax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, (int)(cols / 4), ld_x, amax);
  return check_launch("amax");
}

extern "C" int agnn_split_f16(const float* x, int64_t rows, int64_t cols, int64_t ld_x, const float* amax, void* hi,
                              void* lo, int64_t ld_out, agnn_stream_t stream) {
  return agnn_split_f16_dropout(x, rows, cols, ld_x, amax, hi, lo, ld_out, 0.f, nullptr, 0, stream);
}

extern "C" int agnn_split_f16_dropout(const float* x, int64_t rows, int64_t cols, int64_t ld_x, const float* amax,
                                      void* hi, void* lo, int64_t ld_out, float dropout_p, const uint64_t* rng_state,
                                      uint32_t rng_stream, agnn_stream_t stream) {
  return agnn_split_f16_shifted(x, rows, cols, ld_x, amax, hi, lo, ld_out, dropout_p, rng_state, rng_stream, 1, 0, stream);
}

extern "C" int agnn_split_f16_shifted(const float* x, int64_t rows, int64_t cols, int64_t ld_x, const float* amax,
                                      void* hi, void* lo, int64_t ld_out, float dropout_p, const uint64_t* rng_state,
                                      uint32_t rng_stream, int32_t seq_len, int32_t shift, agnn_stream_t stream) {
  if (rows < 0 || cols < 0 || !amax || !hi || !lo) return fail(AGNN_ERR_ARG, "split_f16: bad arguments");
  if (seq_len < 1) return fail(AGNN_ERR_ARG, "split_f16: sequence length must be positive");
  if (dropout_p < 0.f || dropout_p >= 1.f) return fail(AGNN_ERR_ARG, "split_f16: dropout needs 0 <= p < 1");
  if (rows == 0 || cols == 0) return AGNN_OK;
  if (cols % 4 || (ld_x * 4) % 16 || (ld_out * 2) % 16 || !aligned16(x) || !aligned16(hi) || !aligned16(lo))
    return fail(AGNN_ERR_ARG, "split_f16: needs 16-byte aligned rows and a column count multiple of 4");
  int64_t blocks = ceil_div(rows * (cols / 4), 256);
  if (blocks > kNumSM * 16) blocks = kNumSM * 16;
  split_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, (int)(cols / 4), ld_x, amax,
                                                                        static_cast<__half*>(hi), static_cast<__half*>(lo),
                                                                        ld_out, dropout_p, rng_state, rng_stream, seq_len,
                                                                        shift);
  return check_launch("split_f16");
}

extern "C" int agnn_gemm_split_k(int precision, int64_t M, int64_t N, int64_t K) {
  const int block_k = (precision == AGNN_GEMM_BF16 || precision == AGNN_GEMM_F16X3) ? 64 : 32;
  const int64_t tiles = ceil_div(M, kBlockM) * ceil_div(N, kBlockN);
  const int64_t kb = ceil_div(K, block_k);
  if (tiles >= kNumSM || kb < 16) return 1;
  // Work items = tiles x splits run in waves of kNumSM persistent CTAs.  Pick the split count with the best
  // (wave occupancy) x (useful share of an item: K blocks against ~6 K-block-times of fill + partial store).
  int64_t s = 1;
  double best = 0.0;
  const int64_t s_max = kb / 8 < kNumSM ? kb / 8 : kNumSM;      // at least 8 K blocks per split
  for (int64_t c = 1; c <= s_max; ++c) {
    const int64_t items = tiles * c, waves = ceil_div(items, (int64_t)kNumSM);
    const double kbs = (double)ceil_div(kb, c);
    const double score = (double)items / (double)(waves * kNumSM) * kbs / (kbs + 6.0);
    if (score > best * 1.02) { best = score; s = c; }            // prefer fewer splits unless clearly better
  }
  return s < 1 ? 1 : (int)s;
}

// split counts for the problems of ONE grouped launch: the group as a whole should fill the machine for about two
// waves, every problem getting work items in proportion to its share of the K blocks; at least 8 K blocks per item
// (pipeline fill) and at most 32 splits (the partials of a tile are added by one CTA).
extern "C" int agnn_gemm_group_split_k(int precision, int n_problems, const int64_t* M, const int64_t* N,
                                       const int64_t* K, int32_t* split_out) {
  if (n_problems < 0 || (n_problems && (!M || !N || !K || !split_out))) return fail(AGNN_ERR_ARG, "gemm_group_split_k: bad arguments");
  const int block_k = (precision == AGNN_GEMM_BF16 || precision == AGNN_GEMM_F16X3) ? 64 : 32;
  double total_work = 0.0;
  int64_t total_tiles = 0;
  for (int i = 0; i < n_problems; ++i) {
    const int64_t tiles = ceil_div(M[i], kBlockM) * ceil_div(N[i], kBlockN), kb = ceil_div(K[i], block_k);
    total_work += (double)tiles * (double)kb;
    total_tiles += tiles;
  }
  const double target = total_work / kNumSM >= 32.0 ? 2.0 * kNumSM : 1.0 * kNumSM;
  for (int i = 0; i < n_problems; ++i) {
    const int64_t tiles = ceil_div(M[i], kBlockM) * ceil_div(N[i], kBlockN), kb = ceil_div(K[i], block_k);
    int64_t s = 1;
    if (tiles > 0 && kb >= 16 && total_tiles < 2 * kNumSM) {
      const double share = total_work > 0.0 ? (double)tiles * (double)kb / total_work : 0.0;
      const int64_t items = (int64_t)(share * target + 0.5);
      s = ceil_div(items > tiles ? items : tiles, tiles);
      const int64_t s_max = kb / 8 < 32 ? kb / 8 : 32;
      if (s > s_max) s = s_max;
      if (s < 1) s = 1;
    }
    split_out[i] = (int32_t)s;
  }
  return AGNN_OK;
}

extern "C" size_t agnn_gemm_workspace(int precision, int64_t M, int64_t N, int64_t K, int split_k) {
  (void)precision; (void)K;
  if (split_k <= 1) return 0;
  // partial tiles are stored whole: rows and columns padded to the 128 x 128 tile grid
  return (size_t)split_k * (size_t)part_rows(M) * (size_t)part_ld(N) * sizeof(float);
}

extern "C" int64_t agnn_gemm_tickets(int64_t M, int64_t N, int split_k) {
  return split_k > 1 ? ceil_div(M, kBlockM) * ceil_div(N, kBlockN) : 0;
}

extern "C" int agnn_gemm(int precision, int a_layout, int b_layout, int64_t M, int64_t N, int64_t K, const void* a_hi,
                         const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb, void* c,
                         int64_t ldc, const float* bias, int flags, int split_k, void* workspace,
                         size_t workspace_bytes, agnn_stream_t stream) {
  return agnn_gemm_scaled(precision, a_layout, b_layout, M, N, K, a_hi, a_lo, lda, nullptr, b_hi, b_lo, ldb, nullptr, c,
                          ldc, bias, flags, split_k, workspace, workspace_bytes, stream);
}

extern "C" int agnn_gemm_scaled(int precision, int a_layout, int b_layout, int64_t M, int64_t N, int64_t K,
                                const void* a_hi, const void* a_lo, int64_t lda, const float* amax_a, const void* b_hi,
                                const void* b_lo, int64_t ldb, const float* amax_b, void* c, int64_t ldc,
                                const float* bias, int flags, int split_k, void* workspace, size_t workspace_bytes,
                                agnn_stream_t stream) {
  agnn_gemm_problem_t q;
  memset(&q, 0, sizeof(q));
  q.M = M; q.N = N; q.K = K;
  q.a_hi = a_hi; q.a_lo = a_lo; q.lda = lda; q.amax_a = amax_a;
  q.b_hi = b_hi; q.b_lo = b_lo; q.ldb = ldb; q.amax_b = amax_b;
  q.c = c; q.ldc = ldc; q.bias = bias; q.flags = flags; q.split_k = split_k;
  q.workspace = workspace; q.workspace_bytes = workspace_bytes;
  return agnn_gemm_grouped(precision, a_layout, b_layout, 1, &q, nullptr, 0, stream);
}

extern "C" int agnn_gemm_grouped(int precision, int a_layout, int b_layout, int n_problems,
                                 const agnn_gemm_problem_t* problems, int32_t* tickets, int64_t n_tickets,
                                 agnn_stream_t stream) {
  if (precision != AGNN_GEMM_TF32X3 && precision != AGNN_GEMM_TF32 && precision != AGNN_GEMM_BF16 &&
      precision != AGNN_GEMM_F16X3)
    return fail(AGNN_ERR_ARG, "gemm: unknown precision mode %d", precision);
  if (n_problems < 0 || n_problems > kMaxGroup || (n_problems && !problems))
    return fail(AGNN_ERR_ARG, "gemm: 0..%d problems per launch, got %d", kMaxGroup, n_problems);
  const bool a_mn = a_layout == AGNN_LAYOUT_MN_MAJOR, b_mn = b_layout == AGNN_LAYOUT_MN_MAJOR;
  const bool three = precision == AGNN_GEMM_TF32X3 || precision == AGNN_GEMM_F16X3;
  static thread_local GemmGroup grp;           // 9 KB: kept off the stack
  grp.n_prob = 0;
  grp.chain_blocks = three ? 2 : (1 << 30);    // 24 MMAs per TMEM chain in both three-product modes
  int64_t items = 0, ticket_off = 0;
  for (int i = 0; i < n_problems; ++i) {
    const agnn_gemm_problem_t& q = problems[i];
    if (q.M == 0 || q.N == 0) continue;        // empty problem: nothing to write
    if (q.K == 0) return fail(AGNN_ERR_ARG, "gemm: K == 0");
    GemmParams& p = grp.prob[grp.n_prob];
    int* t = nullptr;
    const int64_t need = agnn_gemm_tickets(q.M, q.N, q.split_k);
    if (tickets && need) {
      if (ticket_off + need > n_tickets)
        return fail(AGNN_ERR_WORKSPACE, "gemm: %lld ticket counters needed, %lld given", (long long)(ticket_off + need),
                    (long long)n_tickets);
      t = tickets + ticket_off;
    }
    int rc = setup_problem(p, precision, a_mn, b_mn, q, t);
    if (rc) return rc;
    if (p.split_k > 1 && t) ticket_off += need;
    grp.tile_start[grp.n_prob] = (int)items;
    items += (int64_t)p.tiles_m * p.tiles_n * p.split_k;
    if (items >= (1ll << 31)) return fail(AGNN_ERR_ARG, "gemm: too many tiles");
    ++grp.n_prob;
  }
  if (grp.n_prob == 0) return AGNN_OK;
  grp.tile_start[grp.n_prob] = (int)items;
  grp.total_tiles = (int)items;
  const int grid = (int)(items < kNumSM ? items : kNumSM);
  cudaStream_t st = (cudaStream_t)stream;
  grp.sched = (items > grid && sched_dynamic()) ? next_sched_counter(st) : nullptr;   // one round: nothing to balance
  int rc;
  if (precision == AGNN_GEMM_BF16) rc = dispatch_layout<kFmtBF16, 1>(a_mn, b_mn, grp, grid, st);
  else if (precision == AGNN_GEMM_F16X3) rc = dispatch_layout<kFmtF16, 3>(a_mn, b_mn, grp, grid, st);
  else if (precision == AGNN_GEMM_TF32X3) rc = dispatch_layout<kFmtTF32, 3>(a_mn, b_mn, grp, grid, st);
  else rc = dispatch_layout<kFmtTF32, 1>(a_mn, b_mn, grp, grid, st);
  if (rc) return rc;
  // without ticket counters the partials of ALL split problems are reduced by one more launch (fixed order as well)
  static thread_local ReduceGroup red;
  red.n_prob = 0;
  int64_t blocks = 0;
  for (int i = 0; i < grp.n_prob; ++i) {
    const GemmParams& p = grp.prob[i];
    if (p.split_k > 1 && !p.tickets) {
      ReduceGroup::P& r = red.prob[red.n_prob];
      r.part = static_cast<const float*>(p.out); r.split_k = p.split_k; r.split_stride = p.split_stride; r.ld_p = p.ldc;
      r.out = p.c_final; r.ldc = p.ldc_final; r.bias = p.bias; r.flags = p.flags; r.M = p.M; r.N = p.N;
      r.amax_out = p.amax_out;
      red.blk_start[red.n_prob++] = (int)blocks;
      int64_t b = ceil_div((int64_t)p.M * ((p.N + 3) / 4), 256);
      if (b > kNumSM * 4) b = kNumSM * 4;
      blocks += b;
    }
  }
  if (red.n_prob) {
    red.blk_start[red.n_prob] = (int)blocks;
    splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(red);
    rc = check_launch("gemm split-K reduce");
    if (rc) return rc;
  }
  return AGNN_OK;
}
