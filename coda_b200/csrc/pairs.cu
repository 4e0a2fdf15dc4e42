// Hypothetical-update "pairs" and the fused EIG quadrature kernel.
//
// A pair is (item b, class c) with the set Z = {h : p_h(b) = c} of models that would be "hit"
// if b were labeled c (coda.py:150-168).  Three exact reductions of the reference's dense
// (B, C, H, P) iteration space (coda.py:261-279):
//   * Z empty      -> the result depends on c only: one "template" pair per class;
//   * Z = {h'}     -> depends on (c, h') only: H template pairs per class;
//   * |Z| >= 2     -> a "heavy" pair, stored with its H-bit mask.
// Pair ids are grouped by class: [z0 template | H singleton templates | heavy pairs of c].
// Every item keeps a CSR list of the pair ids it touches (one per distinct predicted class).
//
//   pair_count / pair_templates / pair_fill     build the structure from the hard predictions
//   pair_rows     coda.py:267-276 for a tile of 32 same-class pairs: D = exp(Z . dL),
//                 prob = D . G_{z}, normalise (coda.py:114), information gain per pair
//   pair_gain     coda.py:274-276 from cached P(best | hypothetical) rows
#include "common.cuh"

#include <stdlib.h>

// ---------------------------------------------------------------------------------------
// structure build: one warp per item
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_histogram(const uint16_t* __restrict__ hrow, int H, int lane, int* cnt) {
  for (int h = lane; h < H; h += 32) atomicAdd(&cnt[hrow[h]], 1);
  __syncwarp();
}
__device__ __forceinline__ void warp_histogram_clear(const uint16_t* __restrict__ hrow, int H, int lane, int* cnt) {
  __syncwarp();
  for (int h = lane; h < H; h += 32) cnt[hrow[h]] = 0;
  __syncwarp();
}

__global__ void __launch_bounds__(256) k_pair_count(const uint16_t* __restrict__ hard, int H, long long N, int C,
                                                    int32_t* __restrict__ ent_cnt, int32_t* __restrict__ cls_heavy) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* cnt_all = reinterpret_cast<int*>(smem_raw);   // [8][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* cnt = cnt_all + (size_t)warp * C;
  for (int c = lane; c < C; c += 32) cnt[c] = 0;
  __syncwarp();
  for (long long n = (long long)blockIdx.x * 8 + warp; n < N; n += (long long)gridDim.x * 8) {
    const uint16_t* hrow = hard + (size_t)n * H;
    warp_histogram(hrow, H, lane, cnt);
    int distinct = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
      int c = c0 + lane;
      int k = c < C ? cnt[c] : 0;
      distinct += __popc(__ballot_sync(CODA_FULL, k >= 1));
      if (k >= 2) atomicAdd(&cls_heavy[c], 1);
    }
    if (lane == 0) ent_cnt[n] = distinct;
    warp_histogram_clear(hrow, H, lane, cnt);
  }
}

// templates: pid = cls_base[c] + 0 (Z empty), cls_base[c] + 1 + h' (Z = {h'})
__global__ void k_pair_templates(int H, int C, int W, const long long* __restrict__ cls_base,
                                 uint32_t* __restrict__ zmask, uint16_t* __restrict__ pair_cls) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)C * (H + 1)) return;
  const int c = (int)(i / (H + 1)), k = (int)(i % (H + 1));
  const long long pid = cls_base[c] + k;
  pair_cls[pid] = (uint16_t)c;
  for (int w = 0; w < W; ++w) zmask[(size_t)pid * W + w] = (k >= 1 && ((k - 1) >> 5) == w) ? (1u << ((k - 1) & 31)) : 0u;
}

__global__ void __launch_bounds__(256) k_pair_fill(const uint16_t* __restrict__ hard, int H, long long N, int C, int W,
                                                   const long long* __restrict__ ent_off,
                                                   const long long* __restrict__ cls_base,
                                                   int32_t* __restrict__ cls_cursor, int32_t* __restrict__ ent_pair,
                                                   uint16_t* __restrict__ ent_cls, uint32_t* __restrict__ zmask,
                                                   uint16_t* __restrict__ pair_cls, int32_t* __restrict__ pair_item) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* cnt_all = reinterpret_cast<int*>(smem_raw);   // [8][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* cnt = cnt_all + (size_t)warp * C;
  for (int c = lane; c < C; c += 32) cnt[c] = 0;
  __syncwarp();
  for (long long n = (long long)blockIdx.x * 8 + warp; n < N; n += (long long)gridDim.x * 8) {
    const uint16_t* hrow = hard + (size_t)n * H;
    warp_histogram(hrow, H, lane, cnt);
    long long pos = ent_off[n];
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int cl = c0 + lane;
      const int k = cl < C ? cnt[cl] : 0;
      uint32_t present = __ballot_sync(CODA_FULL, k >= 1);
      while (present) {
        const int src = __ffs(present) - 1;
        present &= present - 1;
        const int c = c0 + src;
        const int kc = __shfl_sync(CODA_FULL, k, src);
        // mask words: lane w ends up holding word w  (H <= 1024)
        uint32_t myword = 0;
        for (int w = 0; w < W; ++w) {
          const int h = w * 32 + lane;
          const uint32_t bits = __ballot_sync(CODA_FULL, h < H && hrow[h] == c);
          if (lane == w) myword = bits;
        }
        long long pid;
        if (kc == 1) {
          const uint32_t has = __ballot_sync(CODA_FULL, myword != 0);
          const int wl = __ffs(has) - 1;
          const uint32_t word = __shfl_sync(CODA_FULL, myword, wl);
          const int hp = wl * 32 + (__ffs(word) - 1);
          pid = cls_base[c] + 1 + hp;
        } else {
          int slot = 0;
          if (lane == 0) slot = atomicAdd(&cls_cursor[c], 1);
          slot = __shfl_sync(CODA_FULL, slot, 0);
          pid = cls_base[c] + 1 + H + slot;
          if (lane < W) zmask[(size_t)pid * W + lane] = myword;
          if (lane == 0) {
            pair_cls[pid] = (uint16_t)c;
            pair_item[pid] = (int32_t)n;
          }
        }
        if (lane == 0) {
          ent_pair[pos] = (int32_t)pid;
          ent_cls[pos] = (uint16_t)c;
        }
        ++pos;
      }
    }
    warp_histogram_clear(hrow, H, lane, cnt);
  }
}

extern "C" int coda_b200_pair_count(const uint16_t* hard, int H, int64_t N, int C, int32_t* ent_cnt,
                                    int32_t* cls_heavy, coda_stream_t stream) {
  CODA_CHECK_ARG(hard && ent_cnt && cls_heavy, "pair_count: null pointer");
  CODA_CHECK_ARG(H <= 1024, "pair_count: H=%d > 1024 not supported", H);
  size_t smem = (size_t)8 * C * 4;
  CODA_CHECK_ARG(smem <= 200 * 1024, "pair_count: C=%d too large", C);
  CODA_CUDA_OK(cudaFuncSetAttribute(k_pair_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)min((long long)(N + 7) / 8, (long long)coda_sm_count() * 8);
  k_pair_count<<<grid, 256, smem, as_stream(stream)>>>(hard, H, N, C, ent_cnt, cls_heavy);
  CODA_LAUNCH_OK("k_pair_count");
  return CODA_B200_OK;
}

extern "C" int coda_b200_pair_fill(const uint16_t* hard, int H, int64_t N, int C, const int64_t* ent_off,
                                   const int64_t* cls_base, int32_t* cls_cursor, int32_t* ent_pair,
                                   uint16_t* ent_cls, uint32_t* zmask, uint16_t* pair_cls, int32_t* pair_item,
                                   coda_stream_t stream) {
  CODA_CHECK_ARG(hard && ent_off && cls_base && cls_cursor && ent_pair && ent_cls && zmask && pair_cls && pair_item,
                 "pair_fill: null pointer");
  CODA_CHECK_ARG(H <= 1024, "pair_fill: H=%d > 1024 not supported", H);
  const int W = (H + 31) / 32;
  long long nt = (long long)C * (H + 1);
  k_pair_templates<<<(unsigned)((nt + 255) / 256), 256, 0, as_stream(stream)>>>(
      H, C, W, reinterpret_cast<const long long*>(cls_base), zmask, pair_cls);
  CODA_LAUNCH_OK("k_pair_templates");
  size_t smem = (size_t)8 * C * 4;
  CODA_CUDA_OK(cudaFuncSetAttribute(k_pair_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)min((long long)(N + 7) / 8, (long long)coda_sm_count() * 8);
  k_pair_fill<<<grid, 256, smem, as_stream(stream)>>>(hard, H, N, C, W, reinterpret_cast<const long long*>(ent_off),
                                                      reinterpret_cast<const long long*>(cls_base), cls_cursor,
                                                      ent_pair, ent_cls, zmask, pair_cls, pair_item);
  CODA_LAUNCH_OK("k_pair_fill");
  return CODA_B200_OK;
}

// ---------------------------------------------------------------------------------------
// pair_rows: the fused quadrature for a tile of TPAIR same-class pairs (SIMT fp32).
//   phase A  logD[p][x] = sum_h z[p][h] * dL[c][h][x]      threads <-> quadrature node x
//   phase B  a_k[p][h]  = sum_x G_k[c][x][h] * D[p][x]      threads <-> model h  (k = miss, hit)
//   final    prob = z ? a_1 : a_0 ; normalise over h ; gain = sum_h f(m0) - f(m0 + pi_c (ph - PB_c))
// ---------------------------------------------------------------------------------------
#define TPAIR 32
#define TPAD (TPAIR + 4)   // padded row (floats): 16-byte aligned, conflict-free 128-bit rows
#define NODES 256

struct PairRowsArgs {
  const int4* tiles;          // (class, first pid, count, unused)
  const uint32_t* zmask;      // [npairs][W]
  const float* dL;            // [C][H][P]
  const float* G0T;           // [C][P][Hp]
  const float* G1T;           // [C][P][Hp]
  const float* PB;            // [C][Hp]
  const float* m0;            // [Hp]
  const float* pi_hat;        // [C]
  float* ph_cache;            // [npairs][Hp] or null
  float* gain;                // [npairs]
  uint32_t* flags;
  const long long* sel;       // optional: device-resident {idx, class}; then only that class's tiles run
  const long long* tile_off;  // [C+1] first tile of every class (needed with sel)
  int H, Hp, W;
};

template <int HB>   // models per pass: 32, 64, 128 or 256
__global__ void __launch_bounds__(256) k_pair_rows(PairRowsArgs a, int tile0) {
  constexpr int NG = 256 / HB;        // pair groups working side by side in phase B
  constexpr int PPG = TPAIR / NG;     // pairs per group
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = a.H, Hp = a.Hp, W = a.W;
  float* Ds = reinterpret_cast<float*>(smem_raw);                 // [NODES][TPAD]
  float* zf = Ds + NODES * TPAD;                                  // [Hp][TPAD]   (phase A)
  float* Ps = zf;                                                 // [TPAIR][Hp]  (phase B, aliases zf)
  const size_t zf_floats = (size_t)Hp * TPAD > (size_t)TPAIR * Hp ? (size_t)Hp * TPAD : (size_t)TPAIR * Hp;
  uint32_t* zs = reinterpret_cast<uint32_t*>(zf + zf_floats);     // [TPAIR][W]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.sel) {   // host-free loop: the class to refresh is only known on the device
    const long long t = a.sel[1];
    tile0 = (int)a.tile_off[t];
    if ((long long)blockIdx.x >= a.tile_off[t + 1] - a.tile_off[t]) return;
  }
  const int4 tile = a.tiles[tile0 + blockIdx.x];
  const int c = tile.x, pid0 = tile.y, cnt = tile.z;

  for (int i = tid; i < TPAIR * W; i += 256) {
    int p = i / W, w = i % W;
    zs[i] = p < cnt ? a.zmask[(size_t)(pid0 + p) * W + w] : 0u;
  }
  __syncthreads();
  for (int i = tid; i < Hp * TPAIR; i += 256) {
    int h = i / TPAIR, p = i % TPAIR;
    zf[h * TPAD + p] = (float)((zs[p * W + (h >> 5)] >> (h & 31)) & 1u);
  }
  __syncthreads();

  // ---- phase A -------------------------------------------------------------------------
  {
    float acc[TPAIR];
#pragma unroll
    for (int p = 0; p < TPAIR; ++p) acc[p] = 0.f;
    const float* dl = a.dL + (size_t)c * H * NODES + tid;
    int h = 0;
    for (; h + 4 <= H; h += 4) {
      float d[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) d[k] = __ldg(dl + (size_t)(h + k) * NODES);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4* zr = reinterpret_cast<const float4*>(zf + (h + k) * TPAD);
#pragma unroll
        for (int q = 0; q < TPAIR / 4; ++q) {
          float4 z4 = zr[q];
          acc[4 * q + 0] = fmaf(z4.x, d[k], acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(z4.y, d[k], acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(z4.z, d[k], acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(z4.w, d[k], acc[4 * q + 3]);
        }
      }
    }
    for (; h < H; ++h) {
      float d = __ldg(dl + (size_t)h * NODES);
      const float4* zr = reinterpret_cast<const float4*>(zf + h * TPAD);
#pragma unroll
      for (int q = 0; q < TPAIR / 4; ++q) {
        float4 z4 = zr[q];
        acc[4 * q + 0] = fmaf(z4.x, d, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(z4.y, d, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(z4.z, d, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(z4.w, d, acc[4 * q + 3]);
      }
    }
    float4* drow = reinterpret_cast<float4*>(Ds + tid * TPAD);
#pragma unroll
    for (int q = 0; q < TPAIR / 4; ++q)
      drow[q] = make_float4(expf(acc[4 * q + 0]), expf(acc[4 * q + 1]), expf(acc[4 * q + 2]), expf(acc[4 * q + 3]));
  }
  __syncthreads();   // Ds complete; zf no longer needed (Ps aliases it)

  // ---- phase B -------------------------------------------------------------------------
  {
    const int hl = tid % HB, grp = tid / HB;
    const int pbase = grp * PPG;
    for (int h0 = 0; h0 < Hp; h0 += HB) {
      const int h = h0 + hl;
      float a0[PPG], a1[PPG];
#pragma unroll
      for (int p = 0; p < PPG; ++p) { a0[p] = 0.f; a1[p] = 0.f; }
      if (h < Hp) {
        const float* g0p = a.G0T + (size_t)c * NODES * Hp + h;
        const float* g1p = a.G1T + (size_t)c * NODES * Hp + h;
        for (int x = 0; x < NODES; x += 4) {
          float g0[4], g1[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            g0[k] = __ldg(g0p + (size_t)(x + k) * Hp);
            g1[k] = __ldg(g1p + (size_t)(x + k) * Hp);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4* dr = reinterpret_cast<const float4*>(Ds + (x + k) * TPAD + pbase);
#pragma unroll
            for (int q = 0; q < PPG / 4; ++q) {
              float4 d4 = dr[q];
              a0[4 * q + 0] = fmaf(g0[k], d4.x, a0[4 * q + 0]);
              a0[4 * q + 1] = fmaf(g0[k], d4.y, a0[4 * q + 1]);
              a0[4 * q + 2] = fmaf(g0[k], d4.z, a0[4 * q + 2]);
              a0[4 * q + 3] = fmaf(g0[k], d4.w, a0[4 * q + 3]);
              a1[4 * q + 0] = fmaf(g1[k], d4.x, a1[4 * q + 0]);
              a1[4 * q + 1] = fmaf(g1[k], d4.y, a1[4 * q + 1]);
              a1[4 * q + 2] = fmaf(g1[k], d4.z, a1[4 * q + 2]);
              a1[4 * q + 3] = fmaf(g1[k], d4.w, a1[4 * q + 3]);
            }
          }
        }
        const int wsel = h >> 5, bsel = h & 31;
#pragma unroll
        for (int p = 0; p < PPG; ++p) {
          const uint32_t bit = (zs[(pbase + p) * W + wsel] >> bsel) & 1u;
          Ps[(size_t)(pbase + p) * Hp + h] = bit ? a1[p] : a0[p];
        }
      }
    }
  }
  __syncthreads();

  // ---- normalise + information gain: one warp per pair ---------------------------------
  const bool want_gain = a.gain != nullptr;       // cache-only refresh: m0 / pi_hat may not be final yet
  const float pic = want_gain ? a.pi_hat[c] : 0.f;
  const float* pbrow = a.PB + (size_t)c * Hp;
  uint32_t bad = 0;
  for (int p = warp; p < cnt; p += 8) {
    const float* pr = Ps + (size_t)p * Hp;
    float s = 0.f;
    for (int h = lane; h < H; h += 32) s += pr[h];
    s = warp_sum(s);
    if (!isfinite(s)) bad |= CODA_B200_FLAG_NONFINITE_EIG;
    const float den = fmaxf(s, 1e-30f);                         // coda.py:114
    float g = 0.f;
    float* cache = a.ph_cache ? a.ph_cache + (size_t)(pid0 + p) * Hp : nullptr;
    for (int h = lane; h < Hp; h += 32) {
      float ph = 0.f;
      if (h < H) {
        ph = pr[h] / den;
        if (want_gain) {
          const float m = a.m0[h];
          const float mix = m + pic * (ph - pbrow[h]);          // coda.py:274-275
          g += ent_term(m) - ent_term(mix);                     // coda.py:254, 276
        }
      }
      if (cache) cache[h] = ph;
    }
    if (want_gain) {
      g = warp_sum(g);
      if (lane == 0) a.gain[pid0 + p] = g;
    }
  }
  if (bad) atomicOr(a.flags, bad);
}

static size_t pair_rows_smem(int Hp, int W) {
  size_t zf_floats = (size_t)Hp * TPAD > (size_t)TPAIR * Hp ? (size_t)Hp * TPAD : (size_t)TPAIR * Hp;
  return ((size_t)NODES * TPAD + zf_floats) * 4 + (size_t)TPAIR * W * 4;
}

extern "C" int coda_b200_pair_rows(const int32_t* tiles, int tile_lo, int tile_hi, const uint32_t* zmask,
                                   const float* dL, const float* G0T, const float* G1T, const float* PB,
                                   const float* m0, const float* pi_hat, int H, float* ph_cache, float* gain,
                                   const int64_t* sel, const int64_t* tile_off, uint32_t* flags,
                                   coda_stream_t stream) {
  CODA_CHECK_ARG(tiles && zmask && dL && G0T && G1T && PB && flags, "pair_rows: null pointer");
  CODA_CHECK_ARG((gain && m0 && pi_hat) || (!gain && ph_cache), "pair_rows: need gain (+m0, pi_hat) or ph_cache");
  CODA_CHECK_ARG(H >= 1 && H <= 1024, "pair_rows: H=%d out of range", H);
  if (tile_hi <= tile_lo) return CODA_B200_OK;
  PairRowsArgs a;
  a.tiles = reinterpret_cast<const int4*>(tiles);
  a.zmask = zmask; a.dL = dL; a.G0T = G0T; a.G1T = G1T; a.PB = PB; a.m0 = m0; a.pi_hat = pi_hat;
  a.ph_cache = ph_cache; a.gain = gain; a.flags = flags;
  a.sel = reinterpret_cast<const long long*>(sel);
  a.tile_off = reinterpret_cast<const long long*>(tile_off);
  CODA_CHECK_ARG(!sel || tile_off, "pair_rows: sel needs tile_off");
  a.H = H; a.Hp = (H + 31) / 32 * 32; a.W = a.Hp / 32;
  const size_t smem = pair_rows_smem(a.Hp, a.W);
  CODA_CHECK_ARG(smem <= 227 * 1024, "pair_rows: H=%d needs %zu B shared memory", H, smem);
  const int ntiles = tile_hi - tile_lo;
  cudaStream_t st = as_stream(stream);
#define LAUNCH_PR(HB)                                                                                        \
  do {                                                                                                       \
    CODA_CUDA_OK(cudaFuncSetAttribute(k_pair_rows<HB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_pair_rows<HB><<<ntiles, 256, smem, st>>>(a, tile_lo);                                                  \
  } while (0)
  if (a.Hp <= 32) LAUNCH_PR(32);
  else if (a.Hp <= 64) LAUNCH_PR(64);
  else if (a.Hp <= 128) LAUNCH_PR(128);
  else LAUNCH_PR(256);
#undef LAUNCH_PR
  CODA_LAUNCH_OK("k_pair_rows");
  return CODA_B200_OK;
}

// ---------------------------------------------------------------------------------------
// pair_gain: information gain of every pair from the cached P(best | hypothetical) rows.
// HBM-bound stream over ph_cache: one warp per pair, 128-bit loads, two pairs in flight per warp.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float gain4(const float4 ph, const float4 pb, const float4 m, const float4 fm, float pic) {
  float g = fm.x - ent_term(m.x + pic * (ph.x - pb.x));
  g += fm.y - ent_term(m.y + pic * (ph.y - pb.y));
  g += fm.z - ent_term(m.z + pic * (ph.z - pb.z));
  g += fm.w - ent_term(m.w + pic * (ph.w - pb.w));
  return g;
}

__global__ void __launch_bounds__(256) k_pair_gain(const float* __restrict__ ph_cache,
                                                   const uint16_t* __restrict__ pair_cls, long long npairs, int H,
                                                   int Hp, const float* __restrict__ PB, const float* __restrict__ m0,
                                                   const float* __restrict__ pi_hat, float* __restrict__ gain) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* m0s = reinterpret_cast<float*>(smem_raw);   // [Hp]
  float* fm0 = m0s + Hp;                             // [Hp]  f(m0); padded models carry f(0) so they cancel
  for (int h = threadIdx.x; h < Hp; h += blockDim.x) {
    float m = h < H ? m0[h] : 0.f;
    m0s[h] = m;
    fm0[h] = ent_term(m);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * 8;
  const int nq = Hp >> 7;           // float4 per lane per row, full 128-float groups
  const int rem = Hp & 127;         // Hp is a multiple of 32: remainder handled with a lane mask
  for (long long pid = (long long)blockIdx.x * 8 + warp; pid < npairs; pid += 2 * stride) {
    const long long pid2 = pid + stride;
    const bool has2 = pid2 < npairs;
    const int c1 = pair_cls[pid];
    const int c2 = has2 ? pair_cls[pid2] : c1;
    const float pic1 = pi_hat[c1], pic2 = pi_hat[c2];
    const float4* r1 = reinterpret_cast<const float4*>(ph_cache + (size_t)pid * Hp);
    const float4* r2 = reinterpret_cast<const float4*>(ph_cache + (size_t)(has2 ? pid2 : pid) * Hp);
    const float4* b1 = reinterpret_cast<const float4*>(PB + (size_t)c1 * Hp);
    const float4* b2 = reinterpret_cast<const float4*>(PB + (size_t)c2 * Hp);
    const float4* ms = reinterpret_cast<const float4*>(m0s);
    const float4* fs = reinterpret_cast<const float4*>(fm0);
    float g1 = 0.f, g2 = 0.f;
    int q = 0;
    for (; q < nq; ++q) {
      const int i = q * 32 + lane;
      const float4 a1 = __ldg(r1 + i), a2 = __ldg(r2 + i);
      const float4 p1 = __ldg(b1 + i), p2 = __ldg(b2 + i);
      const float4 m = ms[i], fm = fs[i];
      g1 += gain4(a1, p1, m, fm, pic1);
      g2 += gain4(a2, p2, m, fm, pic2);
    }
    if (rem && lane * 4 < rem) {
      const int i = nq * 32 + lane;
      const float4 a1 = __ldg(r1 + i), a2 = __ldg(r2 + i);
      const float4 p1 = __ldg(b1 + i), p2 = __ldg(b2 + i);
      const float4 m = ms[i], fm = fs[i];
      g1 += gain4(a1, p1, m, fm, pic1);
      g2 += gain4(a2, p2, m, fm, pic2);
    }
    g1 = warp_sum(g1);
    g2 = warp_sum(g2);
    if (lane == 0) {
      gain[pid] = g1;
      if (has2) gain[pid2] = g2;
    }
  }
}

// Fast path for Hp = 128 * NQ: pairs are class-sorted, so the class row PB[c], m0 and f(m0) live in
// registers across a run of pairs and the only traffic is the cached row itself, four pairs in flight.
template <int NQ>
__global__ void __launch_bounds__(256) k_pair_gain_fast(const float* __restrict__ ph_cache,
                                                        const uint16_t* __restrict__ pair_cls, long long npairs,
                                                        int H, const float* __restrict__ PB,
                                                        const float* __restrict__ m0,
                                                        const float* __restrict__ pi_hat, float* __restrict__ gain,
                                                        const long long* __restrict__ sel,
                                                        const long long* __restrict__ cls_base, int cls_host,
                                                        int filter) {
  constexpr int Hp = 128 * NQ;
  // filter: 0 = every pair, 1 = every pair except class t, 2 = only class t; t = sel[1] (device) or cls_host
  long long skip_lo = -1, skip_hi = -1, base0 = 0, total = npairs;
  if (filter) {
    const long long t = sel ? sel[1] : cls_host;
    const long long lo = cls_base[t], hi = cls_base[t + 1];
    if (filter == 1) { skip_lo = lo; skip_hi = hi; }
    else { base0 = lo; total = hi - lo; }
  }
  constexpr int CH = 32;   // pairs per warp chunk
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 m[NQ], fm[NQ], pb[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int h = (q * 32 + lane) * 4;
    float4 v = __ldg(reinterpret_cast<const float4*>(m0) + q * 32 + lane);
    if (h + 0 >= H) v.x = 0.f;
    if (h + 1 >= H) v.y = 0.f;
    if (h + 2 >= H) v.z = 0.f;
    if (h + 3 >= H) v.w = 0.f;
    m[q] = v;
    fm[q] = make_float4(ent_term(v.x), ent_term(v.y), ent_term(v.z), ent_term(v.w));
    pb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int cur = -1;
  float pic = 0.f;
  const long long nchunks = (total + CH - 1) / CH;
  for (long long ch = (long long)blockIdx.x * 8 + warp; ch < nchunks; ch += (long long)gridDim.x * 8) {
    const long long p0 = base0 + ch * CH;
    const long long p1 = min(base0 + total, p0 + CH);
    if (p0 >= skip_lo && p1 <= skip_hi) continue;          // chunk entirely inside the excluded class
    for (long long i = p0; i < p1; i += 4) {
      float4 a[4][NQ];
      int cls4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) cls4[j] = pair_cls[min(i + j, npairs - 1)];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long pid = min(i + j, npairs - 1);
        const float4* r = reinterpret_cast<const float4*>(ph_cache + (size_t)pid * Hp);
#pragma unroll
        for (int q = 0; q < NQ; ++q) a[j][q] = __ldg(r + q * 32 + lane);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long pid = i + j;
        if (pid < p1 && !(pid >= skip_lo && pid < skip_hi)) {
          const int c = cls4[j];
          if (c != cur) {
            cur = c;
            pic = pi_hat[c];
#pragma unroll
            for (int q = 0; q < NQ; ++q) pb[q] = __ldg(reinterpret_cast<const float4*>(PB + (size_t)c * Hp) + q * 32 + lane);
          }
          float g = 0.f;
#pragma unroll
          for (int q = 0; q < NQ; ++q) g += gain4(a[j][q], pb[q], m[q], fm[q], pic);
          g = warp_sum(g);
          if (lane == 0) gain[pid] = g;
        }
      }
    }
  }
}

// Bulk-TMA variant of the gain stream: the cached rows of 32 consecutive pairs are one contiguous 32*Hp*4-byte
// blob, staged into a 3-deep shared-memory ring by cp.async.bulk + mbarrier (one elected thread), so the HBM
// stream runs ahead of the entropy arithmetic without holding rows in registers.  Hp = 128 * NQ.
#define PG_CH 32
#define PG_ST 3
template <int NQ>
__global__ void __launch_bounds__(256) k_pair_gain_tma(const float* __restrict__ ph_cache,
                                                       const uint16_t* __restrict__ pair_cls, long long npairs,
                                                       int H, const float* __restrict__ PB,
                                                       const float* __restrict__ m0,
                                                       const float* __restrict__ pi_hat, float* __restrict__ gain,
                                                       const long long* __restrict__ sel,
                                                       const long long* __restrict__ cls_base, int cls_host,
                                                       int filter) {
  constexpr int Hp = 128 * NQ;
  constexpr uint32_t ROW_B = Hp * 4;
  extern __shared__ __align__(128) unsigned char smem_pg[];
  float* ring = reinterpret_cast<float*>(smem_pg);                               // [PG_ST][PG_CH][Hp]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_pg + (size_t)PG_ST * PG_CH * ROW_B);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long skip_lo = -1, skip_hi = -1, base0 = 0, total = npairs;
  if (filter) {
    const long long t = sel ? sel[1] : cls_host;
    const long long lo = cls_base[t], hi = cls_base[t + 1];
    if (filter == 1) { skip_lo = lo; skip_hi = hi; }
    else { base0 = lo; total = hi - lo; }
  }
  const long long nchunks = (total + PG_CH - 1) / PG_CH;
  if (threadIdx.x == 0) {
    for (int s = 0; s < PG_ST; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  float4 m[NQ], fm[NQ], pb[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int h = (q * 32 + lane) * 4;
    float4 v = __ldg(reinterpret_cast<const float4*>(m0) + q * 32 + lane);
    if (h + 0 >= H) v.x = 0.f;
    if (h + 1 >= H) v.y = 0.f;
    if (h + 2 >= H) v.z = 0.f;
    if (h + 3 >= H) v.w = 0.f;
    m[q] = v;
    fm[q] = make_float4(ent_term(v.x), ent_term(v.y), ent_term(v.z), ent_term(v.w));
    pb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  auto issue = [&](long long ch, int s) {
    const long long p0 = base0 + ch * PG_CH;
    const uint32_t cnt = (uint32_t)min((long long)PG_CH, base0 + total - p0);
    mbar_expect_tx(&full[s], cnt * ROW_B);
    tma_load_1d(ring + (size_t)s * PG_CH * Hp, ph_cache + (size_t)p0 * Hp, cnt * ROW_B, &full[s]);
  };
  const long long first = blockIdx.x, stride = gridDim.x;
  if (threadIdx.x == 0)
    for (int s = 0; s < PG_ST; ++s)
      if (first + s * stride < nchunks) issue(first + s * stride, s);
  int cur = -1;
  float pic = 0.f;
  long long it = 0;
  for (long long ch = first; ch < nchunks; ch += stride, ++it) {
    const int s = (int)(it % PG_ST);
    mbar_wait(&full[s], (uint32_t)((it / PG_ST) & 1));
    const long long p0 = base0 + ch * PG_CH;
    const int cnt = (int)min((long long)PG_CH, base0 + total - p0);
    const float* stage = ring + (size_t)s * PG_CH * Hp;
#pragma unroll
    for (int j = 0; j < PG_CH / 8; ++j) {
      const int r = warp * (PG_CH / 8) + j;
      const long long pid = p0 + r;
      if (r < cnt && !(pid >= skip_lo && pid < skip_hi)) {
        const int c = pair_cls[pid];
        if (c != cur) {
          cur = c;
          pic = pi_hat[c];
#pragma unroll
          for (int q = 0; q < NQ; ++q) pb[q] = __ldg(reinterpret_cast<const float4*>(PB + (size_t)c * Hp) + q * 32 + lane);
        }
        const float4* row = reinterpret_cast<const float4*>(stage + (size_t)r * Hp);
        float g = 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) g += gain4(row[q * 32 + lane], pb[q], m[q], fm[q], pic);
        g = warp_sum(g);
        if (lane == 0) gain[pid] = g;
      }
    }
    __syncthreads();                                       // the stage is free again
    if (threadIdx.x == 0 && ch + PG_ST * stride < nchunks) issue(ch + PG_ST * stride, s);
  }
}

extern "C" int coda_b200_pair_gain(const float* ph_cache, const uint16_t* pair_cls, int64_t npairs, int H,
                                   const float* PB, const float* m0, const float* pi_hat, float* gain,
                                   const int64_t* sel, const int64_t* cls_base, int cls_host, int filter,
                                   coda_stream_t stream) {
  CODA_CHECK_ARG(ph_cache && pair_cls && PB && m0 && pi_hat && gain, "pair_gain: null pointer");
  CODA_CHECK_ARG(filter >= 0 && filter <= 2 && (!filter || cls_base), "pair_gain: bad filter");
  const int Hp = (H + 31) / 32 * 32;
  cudaStream_t st = as_stream(stream);
  const long long* seld = reinterpret_cast<const long long*>(sel);
  const long long* cb = reinterpret_cast<const long long*>(cls_base);
  // measured on B200 (cfg3): 0.94 ms for the TMA ring vs 0.83 ms for the register-prefetch kernel below, so the
  // ring is opt-in (CODA_B200_GAIN_TMA=1)
  static const bool use_tma = [] { const char* e = getenv("CODA_B200_GAIN_TMA"); return e && e[0] == '1'; }();
  if (use_tma && Hp % 128 == 0 && Hp <= 512 && (reinterpret_cast<uintptr_t>(ph_cache) & 15) == 0) {
    const size_t smem_t = (size_t)PG_ST * PG_CH * Hp * 4 + PG_ST * 8;
    int grid = (int)min((long long)(npairs + PG_CH - 1) / PG_CH, (long long)coda_sm_count() * (Hp <= 256 ? 2 : 1));
    if (grid < 1) grid = 1;
#define LAUNCH_PT(NQ)                                                                                              \
  do {                                                                                                             \
    CODA_CUDA_OK(cudaFuncSetAttribute(k_pair_gain_tma<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t)); \
    k_pair_gain_tma<NQ><<<grid, 256, smem_t, st>>>(ph_cache, pair_cls, npairs, H, PB, m0, pi_hat, gain, seld, cb,     \
                                                  cls_host, filter);                                               \
  } while (0)
    if (Hp == 128) LAUNCH_PT(1);
    else if (Hp == 256) LAUNCH_PT(2);
    else if (Hp == 384) LAUNCH_PT(3);
    else LAUNCH_PT(4);
#undef LAUNCH_PT
    CODA_LAUNCH_OK("k_pair_gain_tma");
    return CODA_B200_OK;
  }
  if (Hp % 128 == 0 && Hp <= 512) {
    int grid = (int)min((long long)(npairs + 255) / 256, (long long)coda_sm_count() * 6);
    if (grid < 1) grid = 1;
#define LAUNCH_PG(NQ) k_pair_gain_fast<NQ><<<grid, 256, 0, st>>>(ph_cache, pair_cls, npairs, H, PB, m0, pi_hat, gain, seld, cb, cls_host, filter)
    if (Hp == 128) LAUNCH_PG(1);
    else if (Hp == 256) LAUNCH_PG(2);
    else if (Hp == 384) LAUNCH_PG(3);
    else LAUNCH_PG(4);
#undef LAUNCH_PG
    CODA_LAUNCH_OK("k_pair_gain_fast");
    return CODA_B200_OK;
  }
  if (filter == 1) return CODA_B200_OK;   // generic path: one full pass when called with filter 2 (or 0)
  size_t smem = (size_t)2 * Hp * 4;
  int grid = (int)min((long long)(npairs + 15) / 16, (long long)coda_sm_count() * 8);
  if (grid < 1) grid = 1;
  k_pair_gain<<<grid, 256, smem, st>>>(ph_cache, pair_cls, npairs, H, Hp, PB, m0, pi_hat, gain);
  CODA_LAUNCH_OK("k_pair_gain");
  return CODA_B200_OK;
}
