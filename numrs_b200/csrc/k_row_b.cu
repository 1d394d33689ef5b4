// k_row_b.cu -- instantiates the LAYOUT_ROW FFT pass kernels for log2(N) in {9 10 11}
#include "kernels_inst.cuh"
namespace nrb {
void register_row_b(PassTable &t)
{
    register_size<9, LAYOUT_ROW>(t);
    register_size<10, LAYOUT_ROW>(t);
    register_size<11, LAYOUT_ROW>(t);
}
} // namespace nrb
