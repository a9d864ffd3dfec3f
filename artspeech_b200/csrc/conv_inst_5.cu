// explicit instantiations of the implicit-GEMM convolution launchers (part 6 of 6)
#include "conv_igemm_impl.cuh"

namespace asb {
template int launch_conv<128, 32, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
template int launch_conv<128, 64, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
template int launch_conv<64, 32, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
template int launch_halo_sw<64, 64, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
template int launch_conv<32, 64, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
template int launch_halo_sw<32, 64, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
template int launch_conv<16, 64, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
template int launch_halo<16, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
}  // namespace asb
