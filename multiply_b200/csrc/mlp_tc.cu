// tcgen05 engine (placeholder until the kernel lands): fails loudly.
#include "common.cuh"
namespace mp {
size_t tc_pack_bytes() { return 0; }
int tc_pack(Field& f, Arena& a, cudaStream_t st) { f.tc = nullptr; return 0; }
size_t tc_workspace_bytes(int N) { return 0; }
int tc_sdf_list(const Field&, const float*, const int*, const int*, int, float*, void*, size_t, cudaStream_t) {
  set_error("tcgen05 engine not built"); return -9; }
int tc_shade_list(const Field&, const float*, const int*, const int*, int, const float*, float*, float*, float*,
                  float*, float*, void*, size_t, cudaStream_t) { set_error("tcgen05 engine not built"); return -9; }
int tc_bg(const Field&, const float*, const float*, int, float*, float*, void*, size_t, cudaStream_t) {
  set_error("tcgen05 engine not built"); return -9; }
}
