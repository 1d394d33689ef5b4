// k_big.cu -- instantiates the big-tile (persistent, prefetching) FFT pass kernels of fft_pass2.cuh:
// contiguous lines of 2048 / 4096 / 8192 points and strided lines of 512 / 1024 points with 8192-point tiles
#include "kernels_inst.cuh"
namespace nrb {
void register_big(PassTable &t)
{
    register_size2<11, LAYOUT_ROW>(t);
    register_size2<12, LAYOUT_ROW>(t);
    register_size2<13, LAYOUT_ROW>(t);
    register_size2<9, LAYOUT_COL>(t);
    register_size2<10, LAYOUT_COL>(t);
}
} // namespace nrb
