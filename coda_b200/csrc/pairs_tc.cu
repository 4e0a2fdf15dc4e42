// pair_rows on the 5th-generation tensor cores (tcgen05 + TMEM), for Hp = 32..256 models.
//
// Same arithmetic as k_pair_rows (pairs.cu) for a tile of 128 same-class pairs, cast as two GEMMs:
//
//   phase A   logD[128 x 256] = Z[128 x Hp] . dL_c[Hp x 256]             K = Hp (models)
//             Z in {0,1} is exact in bf16; dL is split into 3 bf16 limbs (24 mantissa bits) -> 3 MMAs
//   epilogue  D = exp(logD)  (TMEM -> registers -> exp -> two bf16 limbs -> shared memory, A-operand layout)
//   phase B   prob_k[128 x Hp] = D[128 x 256] . G_k[256 x Hp],  k = miss, hit    K = 256 (quadrature nodes)
//             D = Dhi + Dlo, G = Ghi + Glo (bf16 limbs); Dhi.Ghi + Dhi.Glo + Dlo.Ghi -> 3 MMAs per table
//   epilogue  thread <-> pair (TMEM lane): select hit/miss column by the pair's mask bit, normalise over
//             models (coda.py:114), information gain (coda.py:274-276), cached row write.
//
// Operands are staged by 1-D bulk TMA (cp.async.bulk, no tensor map): the class tables are stored in HBM
// already in the UMMA "no-swizzle, K-major" core-matrix order (8 rows x 16 bytes per core, see tables.cu), so
// one K-chunk of all limbs is a single contiguous blob.  Accumulators live in TMEM (512 columns: logD / prob
// miss in [0,256), prob hit in [256,512)); one CTA per SM, 128 threads, one elected thread issues TMA + MMA.
#include "common.cuh"

#include <cuda_bf16.h>

namespace {

constexpr int TC_M = 128;        // pairs per tile
constexpr int TC_NODES = 256;    // quadrature nodes
constexpr int KA = 32;           // models per phase-A chunk
constexpr int KB = 16;           // nodes per phase-B chunk
constexpr int STAGES_A = 3;
constexpr int TC_THREADS = 512;   // 16 warps: warp w works on TMEM lane quadrant w % 4 (hardware rule) and column part w / 4
constexpr int TC_PARTS = TC_THREADS / 128;
constexpr int STAGES_B = 3;

struct TcArgs {
  const int4* tiles;             // (class, first pid, count <= 128, unused)
  const uint32_t* zmask;         // [npairs][W]   (work-list order)
  const int32_t* row_of;         // [npairs]      work-list position -> row id
  const __nv_bfloat16* dLb;      // [C][Hp/KA][3][TC_NODES x KA]      core-matrix order, rows = nodes
  const __nv_bfloat16* Gb;       // [C][TC_NODES/KB][4][Hp x KB]      core-matrix order, rows = models
  const float* PB;               // [C][Hp]
  const float* m0;               // [Hp]
  const float* pi_hat;           // [C]
  float* ph_cache;               // [npairs][Hp] or null
  float* gain;                   // [npairs] or null
  uint32_t* flags;
  const long long* sel;
  const long long* tile_off;
  int H, Hp, W;
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor (mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
  // version = 1 [46,48), layout_type = SWIZZLE_NONE (0) [61,64)
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}

__device__ __forceinline__ uint32_t instr_desc_bf16(int n) {
  // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b_format BF16 (1) [7,10) [10,13), K-major both,
  // n_dim = N >> 3 [17,23), m_dim = M >> 4 [24,29)
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// byte offset of element (row r, column k) of an operand tile stored as [k_core][r_core][8 rows][8 bf16]
__device__ __forceinline__ uint32_t core_off(int r, int k, int rows) {
  return (uint32_t)(((k >> 3) * (rows >> 3) + (r >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_pair_rows_tc(TcArgs a, int tile0) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int H = a.H, Hp = a.Hp, W = a.W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;      // TMEM lane quadrant (hardware: warp % 4), column part 0..TC_PARTS-1
  const int row = quad * 32 + lane;                 // this thread's pair within the tile
  if (a.sel) {
    const long long t = a.sel[1];
    tile0 = (int)a.tile_off[t];
    if ((long long)blockIdx.x >= a.tile_off[t + 1] - a.tile_off[t]) return;
  }
  const int4 tile = a.tiles[tile0 + blockIdx.x];
  const int c = tile.x, pid0 = tile.y, cnt = tile.z;

  // ---- shared memory carve-up ----------------------------------------------------------------
  // [0, 64K)           phase A: Z operand                       | phase B: D hi
  // [64K, 208K)        phase A: 3 stages x 48K (3 limbs x 256 x KA bf16)
  // [64K, 128K)                                                 | phase B: D lo
  // [128K, 224K)                                                | phase B: 3 stages x 32K (4 tables x Hp x KB)
  // tail               barriers, TMEM base, m0 / PB rows
  unsigned char* opA = smem;
  unsigned char* stg = smem + 128 * 1024;
  unsigned char* tail = stg + 96 * 1024;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(tail);        // [STAGES_A]
  uint64_t* emptyA = fullA + STAGES_A;                        // [STAGES_A]
  uint64_t* fullB = emptyA + STAGES_A;                        // [STAGES_B]
  uint64_t* emptyB = fullB + STAGES_B;                        // [STAGES_B]
  uint64_t* doneA = emptyB + STAGES_B;
  uint64_t* doneB = doneA + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(doneB + 1);
  float* m0s = reinterpret_cast<float*>(tmem_slot + 2);       // [Hp]
  float* pbs = m0s + Hp;                                      // [Hp]
  float* xsum = reinterpret_cast<float*>(stg);                // [TC_PARTS][128] row-sum parts  (stage region: idle in epilogue B)
  float* xgain = xsum + TC_PARTS * TC_M;                      // [TC_PARTS][128] gain parts
  unsigned char* stgA = smem + 64 * 1024;

  if (tid == 0) {
    for (int i = 0; i < STAGES_A; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < STAGES_B; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    mbar_init(doneA, 1);
    mbar_init(doneB, 1);
    mbar_fence_init();
  }
  if (warp == 0) {   // TMEM: all 512 columns (one CTA per SM)
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const bool want_gain = a.gain != nullptr;
  for (int h = tid; h < Hp; h += TC_THREADS) {
    m0s[h] = (want_gain && h < H) ? a.m0[h] : 0.f;
    pbs[h] = a.PB[(size_t)c * Hp + h];
  }
  // ---- Z operand: row = this thread's pair, K = models, bf16 {0, 1} ----------------------------
#pragma unroll 1
  for (int w = half; w < W; w += TC_PARTS) {       // rolled: one copy of the body instead of eight (instruction cache)
    const uint32_t z = row < cnt ? a.zmask[(size_t)(pid0 + row) * W + w] : 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // 8 models -> one 16-byte core row
      const uint32_t b = z >> (8 * q);
      uint4 v;
      v.x = ((b & 1u) ? 0x3F80u : 0u) | ((b & 2u) ? 0x3F800000u : 0u);
      v.y = ((b & 4u) ? 0x3F80u : 0u) | ((b & 8u) ? 0x3F800000u : 0u);
      v.z = ((b & 16u) ? 0x3F80u : 0u) | ((b & 32u) ? 0x3F800000u : 0u);
      v.w = ((b & 64u) ? 0x3F80u : 0u) | ((b & 128u) ? 0x3F800000u : 0u);
      *reinterpret_cast<uint4*>(opA + core_off(row, w * 32 + q * 8, TC_M)) = v;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t opA_s = smem_u32(opA), stg_s = smem_u32(stg), stgA_s = smem_u32(stgA);
  const int nka = Hp / KA;                       // phase-A chunks
  constexpr int nkb = TC_NODES / KB;             // phase-B chunks
  const uint32_t bytesA = 3u * TC_NODES * KA * 2u;          // per stage
  const uint32_t tabB = (uint32_t)Hp * KB * 2u;             // one table, one chunk
  const uint32_t bytesB = 4u * tabB;

  // ---- phase A: logD = Z . dL (3 limbs) --------------------------------------------------------
  // Two elected threads: tid 32 streams the operand chunks (it only ever waits for a stage to be released), tid 0 issues
  // the MMAs (it only ever waits for a stage to be filled) -- so the tensor pipe is never held up by a TMA issue
  // waiting for the previous chunk's MMAs to retire, and all STAGES_A stages are in flight.
  if (tid == 32) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.dLb) + (size_t)c * nka * bytesA;
    for (int kc = 0; kc < nka; ++kc) {
      const int s = kc % STAGES_A;
      if (kc >= STAGES_A) mbar_wait(&emptyA[s], ((kc / STAGES_A) - 1) & 1);
      mbar_expect_tx(&fullA[s], bytesA);
      tma_load_1d(stgA + (size_t)s * 48 * 1024, src + (size_t)kc * bytesA, bytesA, &fullA[s]);
    }
    // prefetch the first phase-B stages while the epilogue below runs (their smem region is free once the
    // phase-A MMAs have completed, which doneA certifies)
    mbar_wait(doneA, 0);
    const unsigned char* srcB = reinterpret_cast<const unsigned char*>(a.Gb) + (size_t)c * nkb * bytesB;
    for (int kc = 0; kc < STAGES_B && kc < nkb; ++kc) {
      mbar_expect_tx(&fullB[kc], bytesB);
      tma_load_1d(stg + (size_t)kc * 32 * 1024, srcB + (size_t)kc * bytesB, bytesB, &fullB[kc]);
    }
  }
  if (tid == 0) {
    const uint32_t idesc = instr_desc_bf16(TC_NODES);
    for (int kc = 0; kc < nka; ++kc) {
      const int s = kc % STAGES_A;
      mbar_wait(&fullA[s], (kc / STAGES_A) & 1);
      tc_fence_after();
#pragma unroll
      for (int limb = 0; limb < 3; ++limb) {
#pragma unroll
        for (int ks = 0; ks < KA / 16; ++ks) {
          // A: Z tile [k_core][16 r_core]: LBO = 16 * 128, advance 2 k-cores per K=16 step
          const uint64_t ad = smem_desc(opA_s + (uint32_t)(kc * (KA / 8) + ks * 2) * (TC_M / 8) * 128, (TC_M / 8) * 128, 128);
          // B: limb tile [4 k_core][32 r_core]: LBO = 32 * 128
          const uint64_t bd = smem_desc(stgA_s + (uint32_t)s * 48 * 1024 + (uint32_t)limb * (TC_NODES * KA * 2) +
                                            (uint32_t)(ks * 2) * (TC_NODES / 8) * 128,
                                        (TC_NODES / 8) * 128, 128);
          mma_bf16(tmem, ad, bd, idesc, (kc | limb | ks) ? 1u : 0u);
        }
      }
      mma_commit(&emptyA[s]);
    }
    mma_commit(doneA);
  }
  __syncwarp();
  mbar_wait(doneA, 0);
  tc_fence_after();

  // ---- epilogue A: D = exp(logD) -> bf16 hi / lo limbs in the A-operand layout ---------------------
  {
    unsigned char* dhi = opA;
    unsigned char* dlo = opA + 64 * 1024;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int cc = half * (TC_NODES / 32 / TC_PARTS); cc < (half + 1) * (TC_NODES / 32 / TC_PARTS); ++cc) {
      float v[32];
      tmem_ld32(trow + cc * 32, v);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d0 = __expf(v[g * 8 + 2 * e]), d1 = __expf(v[g * 8 + 2 * e + 1]);   // ex2.approx: 2 ulp, far inside the 16-bit limb pair D is cut into
          const __nv_bfloat16 h0 = __float2bfloat16_rn(d0), h1 = __float2bfloat16_rn(d1);
          hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lo[e] = pack_bf16(d0 - __bfloat162float(h0), d1 - __bfloat162float(h1));
        }
        const uint32_t off = core_off(row, cc * 32 + g * 8, TC_M);
        *reinterpret_cast<uint4*>(dhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dlo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- phase B: prob_k = D . G_k -----------------------------------------------------------------
  if (tid == 32) {                                // the chunks beyond the STAGES_B prefetched above
    const unsigned char* srcB = reinterpret_cast<const unsigned char*>(a.Gb) + (size_t)c * nkb * bytesB;
    for (int kc = STAGES_B; kc < nkb; ++kc) {
      const int s = kc % STAGES_B;
      mbar_wait(&emptyB[s], ((kc / STAGES_B) - 1) & 1);
      mbar_expect_tx(&fullB[s], bytesB);
      tma_load_1d(stg + (size_t)s * 32 * 1024, srcB + (size_t)kc * bytesB, bytesB, &fullB[s]);
    }
  }
  if (tid == 0) {
    const uint32_t idesc = instr_desc_bf16(Hp);
    const uint32_t lboB = (uint32_t)(Hp / 8) * 128;
    for (int kc = 0; kc < nkb; ++kc) {
      const int s = kc % STAGES_B;
      mbar_wait(&fullB[s], (kc / STAGES_B) & 1);
      tc_fence_after();
      const uint64_t ahi = smem_desc(opA_s + (uint32_t)(kc * 2) * (TC_M / 8) * 128, (TC_M / 8) * 128, 128);
      const uint64_t alo = smem_desc(opA_s + 64 * 1024 + (uint32_t)(kc * 2) * (TC_M / 8) * 128, (TC_M / 8) * 128, 128);
      const uint32_t sb = stg_s + (uint32_t)s * 32 * 1024;
#pragma unroll
      for (int k = 0; k < 2; ++k) {              // k = 0 miss table (G0), 1 hit table (G1)
        const uint64_t ghi = smem_desc(sb + (uint32_t)(2 * k) * tabB, lboB, 128);
        const uint64_t glo = smem_desc(sb + (uint32_t)(2 * k + 1) * tabB, lboB, 128);
        const uint32_t acc = tmem + (uint32_t)k * 256;
        mma_bf16(acc, ahi, ghi, idesc, kc ? 1u : 0u);
        mma_bf16(acc, ahi, glo, idesc, 1u);
        mma_bf16(acc, alo, ghi, idesc, 1u);
      }
      mma_commit(&emptyB[s]);
    }
    mma_commit(doneB);
  }
  __syncwarp();
  mbar_wait(doneB, 0);
  tc_fence_after();

  // ---- epilogue B: (pair, column half) per thread ---------------------------------------------------
  {
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    const int nch = Hp / 32;
    const int per_part = (nch + TC_PARTS - 1) / TC_PARTS;
    const int ch_lo = min(nch, half * per_part), ch_hi = min(nch, (half + 1) * per_part);
    // The chunk loops stay ROLLED: unrolled eight ways (one copy per possible chunk, each warp running two of them) the
    // epilogue was 10 000 SASS lines = 160 KB, and ncu attributed 43 % of the kernel to it with instruction-fetch stalls
    // (stall_no_inst) on top.  The mask words of this thread's (at most two) chunks are re-read (L1 hits).
    uint32_t zsel[2] = {0u, 0u};
    if (row < cnt) {
      if (ch_lo < ch_hi) zsel[0] = a.zmask[(size_t)(pid0 + row) * W + ch_lo];
      if (ch_lo + 1 < ch_hi) zsel[1] = a.zmask[(size_t)(pid0 + row) * W + ch_lo + 1];
    }
    float sum = 0.f;
#pragma unroll 1
    for (int ch = ch_lo; ch < ch_hi; ++ch) {
      float p0[32], p1[32];
      tmem_ld32(trow + ch * 32, p0);
      tmem_ld32(trow + 256 + ch * 32, p1);
      const uint32_t zb = (ch == ch_lo) ? zsel[0] : zsel[1];
#pragma unroll
      for (int i = 0; i < 32; ++i) sum += ((zb >> i) & 1u) ? p1[i] : p0[i];
    }
    xsum[half * TC_M + row] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int q = 0; q < TC_PARTS; ++q) sum += xsum[q * TC_M + row];
    uint32_t bad = 0;
    if (row < cnt && half == 0 && !isfinite(sum)) bad = CODA_B200_FLAG_NONFINITE_EIG;
    const float rden = 1.0f / fmaxf(sum, 1e-30f);                    // coda.py:114
    const float pic = want_gain ? a.pi_hat[c] : 0.f;
    float g = 0.f;
    const int orow = row < cnt ? a.row_of[pid0 + row] : 0;
    if (row < cnt && half == 0 && sum < 0.9999e-30f) bad |= CODA_B200_FLAG_ROWSUM_WARN;    // util.py:37-39
    float* cache = (a.ph_cache && row < cnt) ? a.ph_cache + (size_t)orow * Hp : nullptr;
#pragma unroll 1
    for (int ch = ch_lo; ch < ch_hi; ++ch) {
      float p0[32], p1[32];
      tmem_ld32(trow + ch * 32, p0);
      tmem_ld32(trow + 256 + ch * 32, p1);
      const uint32_t zb = (ch == ch_lo) ? zsel[0] : zsel[1];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int h = ch * 32 + i;
        float ph = (((zb >> i) & 1u) ? p1[i] : p0[i]) * rden;
        if (h >= H) ph = 0.f;
        if (ph < -1e-12f) bad |= CODA_B200_FLAG_NEGATIVE_PROB;          // util.py:33-35
        p0[i] = ph;
        if (want_gain && h < H) {
          const float m = m0s[h];
          g += ent_term(m) - ent_term(m + pic * (ph - pbs[h]));      // coda.py:254, 274-276
        }
      }
      if (cache) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(cache + ch * 32 + i) = make_float4(p0[i], p0[i + 1], p0[i + 2], p0[i + 3]);
      }
    }
    if (want_gain) {
      xgain[half * TC_M + row] = g;
      __syncthreads();
      if (half == 0 && row < cnt) {
        float gs = 0.f;
#pragma unroll
        for (int q = 0; q < TC_PARTS; ++q) gs += xgain[q * TC_M + row];
        a.gain[orow] = gs;
      }
    }
    if (bad) atomicOr(a.flags, bad);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace

extern "C" int coda_b200_pair_rows_tc(const int32_t* tiles128, int tile_lo, int tile_hi, const uint32_t* zmask,
                                      const int32_t* row_of, const void* dLb, const void* Gb, const float* PB, const float* m0,
                                      const float* pi_hat, int H, float* ph_cache, float* gain, const int64_t* sel,
                                      const int64_t* tile_off, uint32_t* flags, coda_stream_t stream) {
  CODA_CHECK_ARG(tiles128 && zmask && row_of && dLb && Gb && PB && flags, "pair_rows_tc: null pointer");
  CODA_CHECK_ARG((gain && m0 && pi_hat) || (!gain && ph_cache), "pair_rows_tc: need gain (+m0, pi_hat) or ph_cache");
  CODA_CHECK_ARG(!sel || tile_off, "pair_rows_tc: sel needs tile_off");
  const int Hp = (H + 31) / 32 * 32;
  CODA_CHECK_ARG(H >= 1 && Hp <= 256, "pair_rows_tc: H=%d needs the SIMT kernel", H);
  if (tile_hi <= tile_lo) return CODA_B200_OK;
  TcArgs a;
  a.tiles = reinterpret_cast<const int4*>(tiles128);
  a.zmask = zmask;
  a.row_of = row_of;
  a.dLb = reinterpret_cast<const __nv_bfloat16*>(dLb);
  a.Gb = reinterpret_cast<const __nv_bfloat16*>(Gb);
  a.PB = PB; a.m0 = m0; a.pi_hat = pi_hat; a.ph_cache = ph_cache; a.gain = gain; a.flags = flags;
  a.sel = reinterpret_cast<const long long*>(sel);
  a.tile_off = reinterpret_cast<const long long*>(tile_off);
  a.H = H; a.Hp = Hp; a.W = Hp / 32;
  const size_t smem = (size_t)(128 + 96) * 1024 + 256 + (size_t)2 * Hp * 4;
  CODA_CUDA_OK(cudaFuncSetAttribute(k_pair_rows_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pair_rows_tc<<<tile_hi - tile_lo, TC_THREADS, smem, as_stream(stream)>>>(a, tile_lo);
  CODA_LAUNCH_OK("k_pair_rows_tc");
  return CODA_B200_OK;
}
