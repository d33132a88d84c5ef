// Peer-memory plumbing of the multi-GPU output exchange (include/waternet_b200.h, "Multi-GPU exchange"):
// a buffer exported over CUDA IPC, mapped by every other rank INTO ITS OWN device context, copy-engine pushes into it,
// and the stream memory operations that signal completion -- nothing here launches a kernel or touches a context
// on the peer device.  Everything acts on the calling thread's current device.
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace wn {

// cuStreamWriteValue32 / cuStreamWaitValue32 through the runtime's driver entry points (the library does not link
// libcuda): int f(CUstream, CUdeviceptr, cuuint32_t, unsigned flags); flags 0 = default write / wait-GEQ
typedef int (*StreamValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);

static StreamValue32Fn stream_value_fn(const char* symbol) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  if (cudaGetDriverEntryPoint(symbol, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return (StreamValue32Fn)fn;
}

static int stream_value32(bool write, void* stream, void* addr, uint32_t value) {
  const char* symbol = write ? "cuStreamWriteValue32" : "cuStreamWaitValue32";
  if (!addr || ((uintptr_t)addr & 3)) {
    set_error("%s: address must be a 4-byte aligned device pointer", symbol);
    return WN_E_INVALID;
  }
  static StreamValue32Fn write_fn = stream_value_fn("cuStreamWriteValue32");
  static StreamValue32Fn wait_fn = stream_value_fn("cuStreamWaitValue32");
  StreamValue32Fn fn = write ? write_fn : wait_fn;
  if (!fn) {
    set_error("%s is not available from this driver", symbol);
    return WN_E_UNSUPPORTED;
  }
  int r = fn((cudaStream_t)stream, (unsigned long long)(uintptr_t)addr, value, 0u);
  if (r != 0) {
    set_error("%s failed with CUresult %d", symbol, r);
    return WN_E_CUDA;
  }
  return WN_OK;
}

}  // namespace wn

using namespace wn;

static_assert(sizeof(cudaIpcMemHandle_t) == WN_PEER_HANDLE_BYTES, "WN_PEER_HANDLE_BYTES must be the size of cudaIpcMemHandle_t");

extern "C" {

int wn_stream_write_value32(void* stream, void* addr, uint32_t value) {
  return stream_value32(true, stream, addr, value);
}

int wn_stream_wait_value32(void* stream, void* addr, uint32_t value) {
  return stream_value32(false, stream, addr, value);
}

int wn_peer_alloc(size_t bytes, void** ptr, unsigned char* handle_out) {
  if (!ptr || !handle_out || bytes == 0) {
    set_error("wn_peer_alloc: bad argument");
    return WN_E_INVALID;
  }
  void* p = nullptr;
  WN_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();  // the zeros are in memory before any peer can write
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("wn_peer_alloc: %s", cudaGetErrorString(e));
    cudaFree(p);
    return WN_E_CUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  *ptr = p;
  return WN_OK;
}

int wn_peer_open(const unsigned char* handle, void** ptr) {
  if (!handle || !ptr) {
    set_error("wn_peer_open: bad argument");
    return WN_E_INVALID;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  WN_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return WN_OK;
}

int wn_peer_close(void* ptr) {
  if (!ptr) return WN_OK;
  WN_CUDA(cudaIpcCloseMemHandle(ptr));
  return WN_OK;
}

int wn_peer_free(void* ptr) {
  if (!ptr) return WN_OK;
  WN_CUDA(cudaFree(ptr));
  return WN_OK;
}

int wn_memcpy_async(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return WN_OK;
  if (!dst || !src) {
    set_error("wn_memcpy_async: null pointer");
    return WN_E_INVALID;
  }
  WN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return WN_OK;
}

}  // extern "C"
