// Peer-memory exchange between the N-axis shards, used INSIDE the single-CTA step kernels.
//
// Every shard owns a mailbox (HBM of its own GPU) that all peers can store into over NVLink (CUDA IPC mapping or
// plain peer access).  Layout of a mailbox:
//   flags  u64 [XCH_NCHAN][2 parities][MAX_WORLD sources]                       (offset 0)
//   data   per channel: [2 parities][world sources][slot_bytes(channel)]
// One exchange on channel ch at epoch e (1, 2, 3, ... counted per channel on the device):
//   push   every thread of the CTA copies its part of the payload into slot (e & 1, my rank) of EVERY mailbox,
//          fences at system scope, the CTA synchronises, then one thread per peer releases flag = e;
//   wait   one thread per source spins (acquire, system scope) on the LOCAL mailbox until the flag reaches e.
// A rank can be at most one epoch ahead of a peer that still reads the previous one (it cannot finish epoch e+1
// without that peer's epoch-e+1 contribution, which the peer sends only after it has consumed epoch e), so two
// parities suffice.  A peer that never arrives is reported after 2 s (flag bit, no hang).
#pragma once
#include "common.cuh"

#define XCH_NCHAN 4
#define XCH_REC 0      // arg-max record + candidate hard rows (device loop) / record only (API)
#define XCH_PISUM 1    // marginal column sums
#define XCH_JROW 2     // owner's p_h(idx) row (API add_label)
#define XCH_REPORT 3   // report block with the tie list (API get_next_item_to_label)

struct XchgView {
  int world, rank;
  unsigned char* box[CODA_B200_MAX_WORLD];
  unsigned long long* epoch;             // [2 * XCH_NCHAN]: epoch counters, then nanoseconds spent waiting, per channel
  uint32_t chan_off[XCH_NCHAN];
  uint32_t slot_bytes[XCH_NCHAN];
};

__host__ __device__ inline uint32_t xch_align16(uint32_t b) { return (b + 15u) & ~15u; }

inline void xchg_layout(int world, int H, int C, int rep_words, uint32_t* chan_off, uint32_t* slot_bytes, size_t* total) {
  const uint32_t hrow = xch_align16((uint32_t)H * 2);
  slot_bytes[XCH_REC] = 64 + 2 * hrow;                                   // record (8 x i64) + rows of candidates A and B
  slot_bytes[XCH_PISUM] = xch_align16((uint32_t)C * 8);
  slot_bytes[XCH_JROW] = 16 + hrow;                                       // {owner flag, pad} + row
  slot_bytes[XCH_REPORT] = xch_align16((uint32_t)rep_words * 8);
  size_t off = (size_t)XCH_NCHAN * 2 * CODA_B200_MAX_WORLD * 8;
  for (int ch = 0; ch < XCH_NCHAN; ++ch) {
    chan_off[ch] = (uint32_t)off;
    off += (size_t)2 * world * slot_bytes[ch];
  }
  *total = (off + 255) & ~(size_t)255;
}

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long* xch_flag(const XchgView& x, int box, int ch, int par, int src) {
  return reinterpret_cast<unsigned long long*>(x.box[box]) + ((size_t)(ch * 2 + par) * CODA_B200_MAX_WORLD + src);
}
__device__ __forceinline__ unsigned char* xch_slot(const XchgView& x, int box, int ch, int par, int src) {
  return x.box[box] + x.chan_off[ch] + (size_t)(par * x.world + src) * x.slot_bytes[ch];
}
// the epoch this kernel's exchange on `ch` runs at (the counter is advanced by xch_done)
__device__ __forceinline__ unsigned long long xch_epoch(const XchgView& x, int ch) { return x.epoch[ch] + 1; }

// all threads of the (single) CTA; src: `bytes` (multiple of 16) of local or shared memory, 16-byte aligned
__device__ __forceinline__ void xch_push(const XchgView& x, int ch, unsigned long long ep, const void* src, uint32_t bytes) {
  const int par = (int)(ep & 1);
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  const uint32_t n4 = bytes >> 4;
  for (int p = 0; p < x.world; ++p) {
    uint4* d4 = reinterpret_cast<uint4*>(xch_slot(x, p, ch, par, x.rank));
    for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = s4[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < x.world) st_release_sys(xch_flag(x, threadIdx.x, ch, par, x.rank), ep);
}
// returns (to every thread) false if some peer did not arrive within 2 s.  The time thread 0 spends here (until the
// slowest peer's contribution has landed: exchange latency + skew between the shards) is added to epoch[4 + ch] (ns).
__device__ __forceinline__ bool xch_wait(const XchgView& x, int ch, unsigned long long ep) {
  __shared__ int s_ok;
  const unsigned long long t_enter = globaltimer_ns();
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < x.world) {
    const unsigned long long* f = xch_flag(x, x.rank, ch, (int)(ep & 1), threadIdx.x);
    const unsigned long long t0 = globaltimer_ns();
    bool ok = true;
    while (ld_acquire_sys(f) < ep) {
      if (globaltimer_ns() - t0 > 2000000000ull) { ok = false; break; }
      __nanosleep(64);
    }
    if (!ok) s_ok = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) x.epoch[XCH_NCHAN + ch] += globaltimer_ns() - t_enter;
  return s_ok != 0;
}
__device__ __forceinline__ const unsigned char* xch_data(const XchgView& x, int ch, unsigned long long ep, int src) {
  return xch_slot(x, x.rank, ch, (int)(ep & 1), src);
}
__device__ __forceinline__ void xch_done(const XchgView& x, int ch, unsigned long long ep) {
  if (threadIdx.x == 0) x.epoch[ch] = ep;
}
#endif

// host: C-ABI struct -> kernel argument
int xchg_view_from(const coda_xchg_t* x, XchgView* out);
