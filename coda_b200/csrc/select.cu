// Per-item EIG assembly, candidate arg-max and the isclose tie scan.
//
//   eig_points    coda.py:278   eig[b] = H_before - sum_c pi_hat_xi[b,c] * H_after[b,c]
//                 written as sum_c xi[b,c] * gain(b,c) with gain = H_before - H_after (sum_c xi = 1)
//                 and gain(b,c) = gain0[c] for every class no model predicts (template pair z0).
//                 coda.py:215-219/239 candidate set: unlabeled & non-unanimous, else all unlabeled.
//   select_merge  coda.py:306, 309: global max with first-index-wins over per-shard partials
//   ties          coda.py:307   torch.isclose(q, best, rtol=1e-8, atol=1e-8) evaluated in fp32
//   device_pick   on-device stand-in for oracle(idx) + the tie rule "lowest index" (bench `value` loop)
#include "common.cuh"

struct Best {
  float v;
  long long i;
};
__device__ __forceinline__ void best_update(Best& b, float v, long long i) {
  if (v > b.v || (v == b.v && i < b.i)) { b.v = v; b.i = i; }
}
__device__ __forceinline__ void best_warp(Best& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(CODA_FULL, b.v, o);
    long long oi = __shfl_xor_sync(CODA_FULL, b.i, o);
    best_update(b, ov, oi);
  }
}

#define IDX_NONE 0x7fffffffffffffffLL

// partial record per block: 5 x int64 {bits(vA), iA, cntA, bits(vB), iB}
// One warp per item.  KC = ceil(C / 32) register slots hold the item's U row (KC = 0: generic loop).
template <int KC>
__global__ void __launch_bounds__(256) k_eig_points(const float* __restrict__ U, long long N, int C,
                                                    const long long* __restrict__ ent_off,
                                                    const int32_t* __restrict__ ent_pair,
                                                    const uint16_t* __restrict__ ent_cls,
                                                    const float* __restrict__ gain,
                                                    const long long* __restrict__ cls_base,
                                                    const uint8_t* __restrict__ labeled,
                                                    const uint8_t* __restrict__ disagree, long long n_offset,
                                                    const int2* __restrict__ ell, int ellK,
                                                    float* __restrict__ eig, long long* __restrict__ partials,
                                                    uint32_t* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* g0 = reinterpret_cast<float*>(smem_raw);   // [C] gain of the "no model predicts c" template
  __shared__ float sv[2][8];
  __shared__ long long si[2][8];
  __shared__ long long sc[8];
  for (int c = threadIdx.x; c < C; c += blockDim.x) g0[c] = gain[cls_base[c]];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Best bA{-INFINITY, IDX_NONE}, bB{-INFINITY, IDX_NONE};
  long long cntA = 0;
  uint32_t bad = 0;
  // IT items per warp iteration: all row / CSR-bound loads first, then the entry loads, then the gain
  // gathers, so each dependent level costs one memory latency for IT items instead of one.
  constexpr int IT = 4;
  constexpr int KR = KC > 0 ? KC : 1;
  const long long stride = (long long)gridDim.x * 8;
  for (long long nb = ((long long)blockIdx.x * 8 + warp) * IT; nb < N; nb += stride * IT) {
    float u[IT][KR];
    long long o0[IT], o1[IT];
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      const long long n = min(nb + i, N - 1);
      if (KC > 0) {
#pragma unroll
        for (int k = 0; k < KR; ++k) {
          const int c = lane + 32 * k;
          u[i][k] = c < C ? __ldg(U + (size_t)n * C + c) : 0.f;
        }
      }
      if (!ell) {
        o0[i] = __ldg(ent_off + n);
        o1[i] = __ldg(ent_off + n + 1);
      } else {
        o0[i] = 0; o1[i] = 0;
      }
    }
    int ec[IT], ep[IT];
    if (ell) {
      // ELL copy of the CSR lists ([N][ellK] x {pair id, class}, -1 padded, ellK >= longest list): the entry
      // address follows from the item index, so these loads are issued together with the U rows above
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        const long long n = min(nb + i, N - 1);
        int2 e2 = make_int2(0, -1);
        if (lane < ellK) e2 = __ldg(ell + (size_t)n * ellK + lane);
        ep[i] = e2.x; ec[i] = e2.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        ec[i] = -1; ep[i] = 0;
        if (o0[i] + lane < o1[i]) { ec[i] = ent_cls[o0[i] + lane]; ep[i] = ent_pair[o0[i] + lane]; }
      }
    }
    float eg[IT], eu[IT];
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      const long long n = min(nb + i, N - 1);
      eg[i] = 0.f; eu[i] = 0.f;
      if (ec[i] >= 0) { eg[i] = __ldg(gain + ep[i]); eu[i] = __ldg(U + (size_t)n * C + ec[i]); }
    }
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      const long long n = nb + i;
      if (n >= N) break;
      const float* urow = U + (size_t)n * C;
      float s = 0.f, e = 0.f;
      if (KC > 0) {
#pragma unroll
        for (int k = 0; k < KR; ++k) {
          const int c = lane + 32 * k;
          s += u[i][k];
          if (c < C) e = fmaf(u[i][k], g0[c], e);
        }
      } else {
        for (int c = lane; c < C; c += 32) {
          const float v = __ldg(urow + c);
          s += v;
          e = fmaf(v, g0[c], e);
        }
      }
      if (ec[i] >= 0) e = fmaf(eu[i], eg[i] - g0[ec[i]], e);
      for (long long k0 = o0[i] + lane + 32; k0 < o1[i]; k0 += 32) {   // > 32 distinct predicted classes
        const int c = ent_cls[k0];
        e = fmaf(__ldg(urow + c), __ldg(gain + ent_pair[k0]) - g0[c], e);
      }
      s = warp_sum(s);
      e = warp_sum(e);
      if (lane == 0) {
        // eig = sum_c xi_c * gain_c with xi = U / max(sum U, 1e-12) (coda.py:230, 278): one division per item
        const float v = e / fmaxf(s, 1e-12f);
        eig[n] = v;
        if (!isfinite(v)) bad |= CODA_B200_FLAG_NONFINITE_EIG;
        if (!labeled[n]) {
          best_update(bB, v, n_offset + n);
          if (disagree[n]) {
            best_update(bA, v, n_offset + n);
            ++cntA;
          }
        }
      }
    }
  }
  if (lane == 0) {
    sv[0][warp] = bA.v; si[0][warp] = bA.i; sv[1][warp] = bB.v; si[1][warp] = bB.i; sc[warp] = cntA;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best a{-INFINITY, IDX_NONE}, b{-INFINITY, IDX_NONE};
    long long cn = 0;
    for (int w = 0; w < 8; ++w) {
      best_update(a, sv[0][w], si[0][w]);
      best_update(b, sv[1][w], si[1][w]);
      cn += sc[w];
    }
    long long* out = partials + (size_t)blockIdx.x * 5;
    out[0] = (long long)__float_as_int(a.v); out[1] = a.i; out[2] = cn;
    out[3] = (long long)__float_as_int(b.v); out[4] = b.i;
  }
  if (bad) atomicOr(flags, bad);
}

// 8 lanes per item (4 items per warp instruction, IT8 batches in flight): for C <= 128 the per-item fixed cost
// (shuffle reductions, address arithmetic) of the warp-per-item kernel above dominates, so narrower groups
// halve the issued instructions per item.  Needs the ELL lists (ellK <= 32).
template <int KC8>
__global__ void __launch_bounds__(256) k_eig_points_g8(const float* __restrict__ U, long long N, int C,
                                                       const float* __restrict__ gain,
                                                       const long long* __restrict__ cls_base,
                                                       const uint8_t* __restrict__ labeled,
                                                       const uint8_t* __restrict__ disagree, long long n_offset,
                                                       const int2* __restrict__ ell, int ellK,
                                                       float* __restrict__ eig, long long* __restrict__ partials,
                                                       uint32_t* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* g0 = reinterpret_cast<float*>(smem_raw);   // [C]
  __shared__ float sv[2][8];
  __shared__ long long si[2][8];
  __shared__ long long sc[8];
  for (int c = threadIdx.x; c < C; c += blockDim.x) g0[c] = gain[cls_base[c]];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane & 7, grp = lane >> 3;
  Best bA{-INFINITY, IDX_NONE}, bB{-INFINITY, IDX_NONE};
  long long cntA = 0;
  uint32_t bad = 0;
  constexpr int IT8 = 2;
  const long long per_iter = (long long)gridDim.x * 8 * 4 * IT8;
  // the loop bound is warp-uniform (full-mask shuffles inside); a group past the end clamps its loads and skips its writes
  for (long long wb = ((long long)blockIdx.x * 8 + warp) * 4 * IT8; wb < N; wb += per_iter) {
    const long long nb = wb + grp * IT8;
    float u[IT8][KC8];
    int2 en[IT8][4];
#pragma unroll
    for (int i = 0; i < IT8; ++i) {
      const long long n = min(nb + i, N - 1);
      const float* urow = U + (size_t)n * C;
#pragma unroll
      for (int k = 0; k < KC8; ++k) {
        const int c = g + 8 * k;
        u[i][k] = c < C ? __ldg(urow + c) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = g + 8 * j;
        en[i][j] = e < ellK ? __ldg(ell + (size_t)n * ellK + e) : make_int2(0, -1);
      }
    }
    float eg[IT8][4], eu[IT8][4];
#pragma unroll
    for (int i = 0; i < IT8; ++i) {
      const long long n = min(nb + i, N - 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        eg[i][j] = 0.f; eu[i][j] = 0.f;
        if (en[i][j].y >= 0) {
          eg[i][j] = __ldg(gain + en[i][j].x);
          eu[i][j] = __ldg(U + (size_t)n * C + en[i][j].y);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < IT8; ++i) {
      const long long n = nb + i;
      float s = 0.f, e = 0.f;
#pragma unroll
      for (int k = 0; k < KC8; ++k) {
        const int c = g + 8 * k;
        s += u[i][k];
        if (c < C) e = fmaf(u[i][k], g0[c], e);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (en[i][j].y >= 0) e = fmaf(eu[i][j], eg[i][j] - g0[en[i][j].y], e);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        s += __shfl_xor_sync(CODA_FULL, s, o);
        e += __shfl_xor_sync(CODA_FULL, e, o);
      }
      if (g == 0 && n < N) {
        const float v = e / fmaxf(s, 1e-12f);                // coda.py:230, 278
        eig[n] = v;
        if (!isfinite(v)) bad |= CODA_B200_FLAG_NONFINITE_EIG;
        if (!labeled[n]) {
          best_update(bB, v, n_offset + n);
          if (disagree[n]) {
            best_update(bA, v, n_offset + n);
            ++cntA;
          }
        }
      }
    }
  }
  best_warp(bA);
  best_warp(bB);
  cntA = warp_sum(cntA);
  if (lane == 0) {
    sv[0][warp] = bA.v; si[0][warp] = bA.i; sv[1][warp] = bB.v; si[1][warp] = bB.i; sc[warp] = cntA;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best a{-INFINITY, IDX_NONE}, b{-INFINITY, IDX_NONE};
    long long cn = 0;
    for (int w = 0; w < 8; ++w) {
      best_update(a, sv[0][w], si[0][w]);
      best_update(b, sv[1][w], si[1][w]);
      cn += sc[w];
    }
    long long* out = partials + (size_t)blockIdx.x * 5;
    out[0] = (long long)__float_as_int(a.v); out[1] = a.i; out[2] = cn;
    out[3] = (long long)__float_as_int(b.v); out[4] = b.i;
  }
  if (bad) atomicOr(flags, bad);
}

extern "C" int coda_b200_eig_blocks(int64_t N) {
  long long want = (N + 31) / 32;
  long long cap = (long long)coda_sm_count() * 8;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

extern "C" int coda_b200_eig_points(const float* U, int64_t N, int C, const int64_t* ent_off, const int32_t* ent_pair,
                                    const uint16_t* ent_cls, const float* gain, const int64_t* cls_base,
                                    const uint8_t* labeled, const uint8_t* disagree, int64_t n_offset,
                                    const int32_t* ell, int ell_k, float* eig, int64_t* partials, uint32_t* flags,
                                    coda_stream_t stream) {
  CODA_CHECK_ARG(U && ent_off && ent_pair && ent_cls && gain && cls_base && labeled && disagree && eig && partials && flags,
                 "eig_points: null pointer");
  size_t smem = (size_t)C * 4;
  CODA_CHECK_ARG(smem <= 48 * 1024, "eig_points: C=%d too large", C);
  CODA_CHECK_ARG(!ell || (ell_k >= 1 && ell_k <= 32), "eig_points: ell_k=%d out of range", ell_k);
  int grid = coda_b200_eig_blocks(N);
  if (ell && C <= 128) {
#define LAUNCH_G8(K8)                                                                                             \
  k_eig_points_g8<K8><<<grid, 256, smem, as_stream(stream)>>>(                                                    \
      U, N, C, gain, reinterpret_cast<const long long*>(cls_base), labeled, disagree, n_offset,                   \
      reinterpret_cast<const int2*>(ell), ell_k, eig, reinterpret_cast<long long*>(partials), flags)
    if (C <= 32) LAUNCH_G8(4);
    else if (C <= 64) LAUNCH_G8(8);
    else if (C <= 104) LAUNCH_G8(13);
    else LAUNCH_G8(16);
#undef LAUNCH_G8
    CODA_LAUNCH_OK("k_eig_points_g8");
    return CODA_B200_OK;
  }
#define LAUNCH_EP(KC)                                                                                              \
  k_eig_points<KC><<<grid, 256, smem, as_stream(stream)>>>(                                                        \
      U, N, C, reinterpret_cast<const long long*>(ent_off), ent_pair, ent_cls, gain,                               \
      reinterpret_cast<const long long*>(cls_base), labeled, disagree, n_offset,                                   \
      reinterpret_cast<const int2*>(ell), ell_k, eig, reinterpret_cast<long long*>(partials), flags)
  if (C <= 32) LAUNCH_EP(1);
  else if (C <= 64) LAUNCH_EP(2);
  else if (C <= 128) LAUNCH_EP(4);
  else LAUNCH_EP(0);
#undef LAUNCH_EP
  CODA_LAUNCH_OK("k_eig_points");
  return CODA_B200_OK;
}

// merge `nrec` partial records (from this shard's blocks, or one per rank after an all-gather)
// into one record {bits(vA), iA, cntA, bits(vB), iB}.  nrec is small (<= a few thousand): one block.
__global__ void __launch_bounds__(256) k_select_merge(const long long* __restrict__ recs, int nrec,
                                                      long long* __restrict__ out) {
  __shared__ float sv[2][8];
  __shared__ long long si[2][8];
  __shared__ long long sc[8];
  Best a{-INFINITY, IDX_NONE}, b{-INFINITY, IDX_NONE};
  long long cn = 0;
  for (int r = threadIdx.x; r < nrec; r += blockDim.x) {
    const long long* rec = recs + (size_t)r * 5;
    best_update(a, __int_as_float((int)rec[0]), rec[1]);
    best_update(b, __int_as_float((int)rec[3]), rec[4]);
    cn += rec[2];
  }
  best_warp(a);
  best_warp(b);
  cn = warp_sum(cn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[0][warp] = a.v; si[0][warp] = a.i; sv[1][warp] = b.v; si[1][warp] = b.i; sc[warp] = cn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best fa{-INFINITY, IDX_NONE}, fb{-INFINITY, IDX_NONE};
    long long c = 0;
    for (int w = 0; w < 8; ++w) {
      best_update(fa, sv[0][w], si[0][w]);
      best_update(fb, sv[1][w], si[1][w]);
      c += sc[w];
    }
    out[0] = (long long)__float_as_int(fa.v); out[1] = fa.i; out[2] = c;
    out[3] = (long long)__float_as_int(fb.v); out[4] = fb.i;
  }
}

extern "C" int coda_b200_select_merge(const int64_t* recs, int nrec, int64_t* out, coda_stream_t stream) {
  CODA_CHECK_ARG(recs && out && nrec >= 1, "select_merge: bad arguments");
  k_select_merge<<<1, 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(recs), nrec,
                                                   reinterpret_cast<long long*>(out));
  CODA_LAUNCH_OK("k_select_merge");
  return CODA_B200_OK;
}

// tie scan against the GLOBAL record `best` (after select_merge over all shards).
// tie_hdr: {count, min tied global index}; tie_idx/tie_val hold up to `cap` entries (unordered).
__global__ void __launch_bounds__(256) k_ties(const float* __restrict__ eig, long long N,
                                              const uint8_t* __restrict__ labeled,
                                              const uint8_t* __restrict__ disagree, long long n_offset,
                                              const long long* __restrict__ best, int cap,
                                              long long* __restrict__ tie_hdr, long long* __restrict__ tie_idx,
                                              float* __restrict__ tie_val) {
  const bool useA = best[2] > 0;                                      // coda.py:239 `or` fallback
  const float bv = __int_as_float((int)(useA ? best[0] : best[3]));
  const float tol = 1e-8f + fabsf(1e-8f * bv);                        // atol + rtol * |best| (fp32)
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (long long)gridDim.x * blockDim.x) {
    if (labeled[n]) continue;
    if (useA && !disagree[n]) continue;
    const float e = eig[n];
    if (e == bv || fabsf(e - bv) <= tol) {
      const long long g = n_offset + n;
      unsigned long long k = atomicAdd(reinterpret_cast<unsigned long long*>(tie_hdr), 1ull);
      atomicMin(tie_hdr + 1, g);
      if (k < (unsigned long long)cap) {
        tie_idx[k] = g;
        tie_val[k] = e;
      }
    }
  }
}

__global__ void k_ties_reset(long long* tie_hdr) {
  tie_hdr[0] = 0;
  tie_hdr[1] = IDX_NONE;
}

extern "C" int coda_b200_ties(const float* eig, int64_t N, const uint8_t* labeled, const uint8_t* disagree,
                              int64_t n_offset, const int64_t* best, int cap, int64_t* tie_hdr, int64_t* tie_idx,
                              float* tie_val, coda_stream_t stream) {
  CODA_CHECK_ARG(eig && labeled && disagree && best && tie_hdr && tie_idx && tie_val, "ties: null pointer");
  int grid = (int)min((long long)(N + 255) / 256, (long long)coda_sm_count() * 4);
  if (grid < 1) grid = 1;
  k_ties_reset<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<long long*>(tie_hdr));
  k_ties<<<grid, 256, 0, as_stream(stream)>>>(eig, N, labeled, disagree, n_offset,
                                              reinterpret_cast<const long long*>(best), cap,
                                              reinterpret_cast<long long*>(tie_hdr),
                                              reinterpret_cast<long long*>(tie_idx), tie_val);
  CODA_LAUNCH_OK("k_ties");
  return CODA_B200_OK;
}

// device-resident stand-in for `oracle(idx)` (coda/oracle.py:23-24): picks the lowest tied global
// index (tie_hdr[1], already min-reduced across shards by the caller), looks the label up in a
// device-resident label vector and writes the {local idx or -1, class} record the update kernels read.
__global__ void k_device_pick(const long long* __restrict__ tie_hdr, const long long* __restrict__ best,
                              const long long* __restrict__ labels_global,
                              long long n_offset, long long N, const float* __restrict__ eig,
                              long long* __restrict__ sel, long long* __restrict__ hist_idx,
                              float* __restrict__ hist_q, long long step) {
  // tie_hdr given: lowest index among the isclose-tied candidates; else the merged arg-max record (first index
  // wins on exactly equal values, torch.argmax; coda.py:309)
  const long long g = tie_hdr ? tie_hdr[1] : (best[2] > 0 ? best[1] : best[4]);
  const long long loc = g - n_offset;
  const bool own = loc >= 0 && loc < N;
  sel[0] = own ? loc : -1;
  sel[1] = labels_global[g];
  if (hist_idx) hist_idx[step] = g;
  if (hist_q) hist_q[step] = own ? eig[loc] : 0.f;
}

extern "C" int coda_b200_device_pick(const int64_t* tie_hdr, const int64_t* best, const int64_t* labels_global,
                                     int64_t n_offset, int64_t N, const float* eig, int64_t* sel, int64_t* hist_idx,
                                     float* hist_q, int64_t step, coda_stream_t stream) {
  CODA_CHECK_ARG((tie_hdr || best) && labels_global && eig && sel, "device_pick: null pointer");
  k_device_pick<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(tie_hdr),
                                                reinterpret_cast<const long long*>(best),
                                                reinterpret_cast<const long long*>(labels_global), n_offset, N, eig,
                                                reinterpret_cast<long long*>(sel),
                                                reinterpret_cast<long long*>(hist_idx), hist_q, step);
  CODA_LAUNCH_OK("k_device_pick");
  return CODA_B200_OK;
}

// ELL copy of the per-item CSR lists (eig_points): ell[n][k] = {pair id, class}, {0, -1} padding.
__global__ void k_ell_build(const long long* __restrict__ ent_off, const int32_t* __restrict__ ent_pair,
                            const uint16_t* __restrict__ ent_cls, long long N, int K, int2* __restrict__ ell) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const long long n = i / K;
  const int k = (int)(i % K);
  const long long o = ent_off[n] + k;
  ell[i] = o < ent_off[n + 1] ? make_int2(ent_pair[o], (int)ent_cls[o]) : make_int2(0, -1);
}

extern "C" int coda_b200_ell_build(const int64_t* ent_off, const int32_t* ent_pair, const uint16_t* ent_cls, int64_t N,
                                   int K, int32_t* ell, coda_stream_t stream) {
  CODA_CHECK_ARG(ent_off && ent_pair && ent_cls && ell && K >= 1 && K <= 32, "ell_build: bad arguments");
  const long long tot = (long long)N * K;
  k_ell_build<<<(unsigned)((tot + 255) / 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const long long*>(ent_off), ent_pair, ent_cls, N, K, reinterpret_cast<int2*>(ell));
  CODA_LAUNCH_OK("k_ell_build");
  return CODA_B200_OK;
}
