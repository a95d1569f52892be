// Instantiations of the masked scan kernel for dt-rank slab count KS = 5 (dt_proj fused on the tensor cores; 16-bit I/O).
#include "masked_scan_kernel.cuh"

namespace actk {
template void launch_ks<__half, 5>(bool, int, dim3, cudaStream_t, const MaskedParams<__half> &, const MaskedMaps &);
template void launch_ks<__nv_bfloat16, 5>(bool, int, dim3, cudaStream_t, const MaskedParams<__nv_bfloat16> &, const MaskedMaps &);
}  // namespace actk
