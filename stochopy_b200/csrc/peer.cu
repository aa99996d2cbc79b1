// Peer mailboxes: allocation and CUDA-IPC mapping (host side of peer.cuh).
// The mailbox is plain cudaMalloc memory (IPC handles cannot be exported from
// stream-ordered / virtual-memory pools), zero-filled so every flag starts at epoch 0.
#include <cstring>

#include "peer.cuh"

using namespace sp;

static_assert(sizeof(cudaIpcMemHandle_t) == SP_PEER_HANDLE_BYTES, "handle size");

extern "C" {

int64_t sp_peer_bytes(int dtype, int world, int64_t ld, int64_t P_total) {
  if (world < 1 || ld < 1 || P_total < 1 || (dtype != SP_F32 && dtype != SP_F64)) return -1;
  return (int64_t)peer_layout(world, ld, P_total, dtype == SP_F32 ? 4 : 8).total;
}

int sp_peer_alloc(int64_t bytes, void** d_mailbox, void* handle_out) {
  SP_CHECK_ARG(bytes > 0 && d_mailbox != nullptr && handle_out != nullptr, "bytes / null pointer");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("sp_peer_alloc: %s", cudaGetErrorString(e));
    if (p) cudaFree(p);
    return SP_ERR_CUDA;
  }
  std::memcpy(handle_out, &h, sizeof(h));
  *d_mailbox = p;
  return SP_OK;
}

int sp_peer_open(const void* handle, void** d_mailbox) {
  SP_CHECK_ARG(handle != nullptr && d_mailbox != nullptr, "null pointer");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("sp_peer_open: %s", cudaGetErrorString(e));
    return SP_ERR_CUDA;
  }
  *d_mailbox = p;
  return SP_OK;
}

int sp_peer_close(void* d_mailbox) {
  SP_CHECK_ARG(d_mailbox != nullptr, "null pointer");
  cudaError_t e = cudaIpcCloseMemHandle(d_mailbox);
  if (e != cudaSuccess) {
    set_error("sp_peer_close: %s", cudaGetErrorString(e));
    return SP_ERR_CUDA;
  }
  return SP_OK;
}

int sp_peer_free(void* d_mailbox) {
  SP_CHECK_ARG(d_mailbox != nullptr, "null pointer");
  cudaError_t e = cudaFree(d_mailbox);
  if (e != cudaSuccess) {
    set_error("sp_peer_free: %s", cudaGetErrorString(e));
    return SP_ERR_CUDA;
  }
  return SP_OK;
}

}  // extern "C"
