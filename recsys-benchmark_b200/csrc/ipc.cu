// Peer-shareable device buffers for the row-sharded tables: plain cudaMalloc allocations
// (so that a CUDA IPC handle maps the exact buffer, independent of torch's caching
// allocator) exported to / imported from the other ranks of the box.  Reads and atomics
// on the imported pointers travel over NVLink / NVSwitch.
#include <string.h>

#include "common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == RSB_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" RSB_API int rsb_shared_alloc(int64_t bytes, void** dev_ptr_out) {
  if (bytes <= 0 || dev_ptr_out == nullptr) return RSB_ERR_BAD_ARG;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, (size_t)bytes);
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  *dev_ptr_out = p;
  return RSB_OK;
}

extern "C" RSB_API int rsb_shared_free(void* dev_ptr) {
  if (dev_ptr == nullptr) return RSB_OK;
  return (int)cudaFree(dev_ptr);
}

extern "C" RSB_API int rsb_ipc_get_handle(const void* dev_ptr, uint8_t* h_handle) {
  if (dev_ptr == nullptr || h_handle == nullptr) return RSB_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr));
  if (e != cudaSuccess) return (int)e;
  memcpy(h_handle, &h, sizeof(h));
  return RSB_OK;
}

extern "C" RSB_API int rsb_ipc_open_handle(const uint8_t* h_handle, void** dev_ptr_out) {
  if (h_handle == nullptr || dev_ptr_out == nullptr) return RSB_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return (int)e;
  *dev_ptr_out = p;
  return RSB_OK;
}

extern "C" RSB_API int rsb_ipc_close_handle(void* dev_ptr) {
  if (dev_ptr == nullptr) return RSB_OK;
  return (int)cudaIpcCloseMemHandle(dev_ptr);
}
