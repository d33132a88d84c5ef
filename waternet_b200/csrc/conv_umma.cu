// placeholder until the tcgen05 path lands
#include "common.cuh"
namespace wn {
int umma_pack_weights(wn_handle*, const float* const*, cudaStream_t) { return WN_OK; }
void umma_free(wn_handle*) {}
size_t umma_forward_workspace_bytes(int n, int h, int w) { return simt_forward_workspace_bytes(n, h, w); }
int umma_forward(wn_handle* h, const float* const in[4], const int64_t in_strides[4][4], float* out,
                 int n, int height, int width, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  return simt_forward(h, in, in_strides, out, n, height, width, workspace, workspace_bytes, stream);
}
}
