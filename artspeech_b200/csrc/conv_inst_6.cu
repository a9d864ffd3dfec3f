// explicit instantiations of the 2-CTA (cta_group::2) implicit-GEMM convolution launcher
#include "conv_igemm_impl.cuh"

namespace asb {
template int launch_conv_2cta<64, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, cudaStream_t);
template int launch_conv_2cta<64, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, cudaStream_t);
}  // namespace asb
