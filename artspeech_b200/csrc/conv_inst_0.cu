// explicit instantiations of the implicit-GEMM convolution launchers (part 1 of 6)
#include "conv_igemm_impl.cuh"

namespace asb {
template int launch_conv<256, 32, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
template int launch_halo_sw<128, 32, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
template int launch_conv<64, 64, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
template int launch_halo<64, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
template int launch_halo<32, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
template int launch_halo_sw<16, 32, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
}  // namespace asb
