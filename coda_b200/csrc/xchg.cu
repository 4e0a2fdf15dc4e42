// Mailbox allocation and mapping for the peer-memory exchange (see xchg.cuh).
#include "xchg.cuh"

int xchg_view_from(const coda_xchg_t* x, XchgView* v) {
  memset(v, 0, sizeof(*v));
  if (!x || x->world <= 1) {
    v->world = 1;
    v->rank = 0;
    return CODA_B200_OK;
  }
  CODA_CHECK_ARG(x->world <= CODA_B200_MAX_WORLD && x->rank >= 0 && x->rank < x->world, "xchg: bad world/rank %d/%d",
                 x->world, x->rank);
  CODA_CHECK_ARG(x->epoch, "xchg: epoch counters missing");
  v->world = x->world;
  v->rank = x->rank;
  for (int p = 0; p < x->world; ++p) {
    CODA_CHECK_ARG(x->box[p], "xchg: mailbox of rank %d missing", p);
    v->box[p] = reinterpret_cast<unsigned char*>(x->box[p]);
  }
  v->epoch = reinterpret_cast<unsigned long long*>(x->epoch);
  size_t total;
  xchg_layout(x->world, x->H, x->C, x->rep_words, v->chan_off, v->slot_bytes, &total);
  return CODA_B200_OK;
}

extern "C" size_t coda_b200_xchg_box_bytes(int world, int H, int C, int rep_words) {
  uint32_t off[XCH_NCHAN], sb[XCH_NCHAN];
  size_t total = 0;
  xchg_layout(world, H, C, rep_words, off, sb, &total);
  return total;
}

extern "C" int coda_b200_xchg_alloc(size_t bytes, void** box_out) {
  CODA_CHECK_ARG(box_out && bytes > 0, "xchg_alloc: bad arguments");
  void* p = nullptr;
  CODA_CUDA_OK(cudaMalloc(&p, bytes));
  CODA_CUDA_OK(cudaMemset(p, 0, bytes));
  CODA_CUDA_OK(cudaDeviceSynchronize());
  *box_out = p;
  return CODA_B200_OK;
}

extern "C" int coda_b200_xchg_free(void* box) {
  if (box) CODA_CUDA_OK(cudaFree(box));
  return CODA_B200_OK;
}

extern "C" int coda_b200_ipc_export(const void* box, void* handle64_host) {
  CODA_CHECK_ARG(box && handle64_host, "ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  CODA_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(box)));
  memcpy(handle64_host, &h, 64);
  return CODA_B200_OK;
}

extern "C" int coda_b200_ipc_open(const void* handle64_host, void** box_out) {
  CODA_CHECK_ARG(handle64_host && box_out, "ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  void* p = nullptr;
  CODA_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *box_out = p;
  return CODA_B200_OK;
}

extern "C" int coda_b200_ipc_close(void* box) {
  if (box) CODA_CUDA_OK(cudaIpcCloseMemHandle(box));
  return CODA_B200_OK;
}

extern "C" int coda_b200_peer_enable(int peer_device) {
  int dev = 0;
  CODA_CUDA_OK(cudaGetDevice(&dev));
  if (dev == peer_device) return CODA_B200_OK;
  int can = 0;
  CODA_CUDA_OK(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  if (!can) {
    coda_set_error("device %d cannot access device %d's memory (no NVLink / P2P path)", dev, peer_device);
    return CODA_B200_ECUDA;
  }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return CODA_B200_OK;
  }
  CODA_CUDA_OK(e);
  return CODA_B200_OK;
}
