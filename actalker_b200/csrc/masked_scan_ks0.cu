// Instantiations of the masked scan kernel for dt-rank slab count KS = 0 (delta tensors from HBM; all I/O dtypes).
#include "masked_scan_kernel.cuh"

namespace actk {
template void launch_ks<float, 0>(bool, int, dim3, cudaStream_t, const MaskedParams<float> &, const MaskedMaps &);
template void launch_ks<__half, 0>(bool, int, dim3, cudaStream_t, const MaskedParams<__half> &, const MaskedMaps &);
template void launch_ks<__nv_bfloat16, 0>(bool, int, dim3, cudaStream_t, const MaskedParams<__nv_bfloat16> &, const MaskedMaps &);
}  // namespace actk
