// k_col_a.cu -- instantiates the LAYOUT_COL FFT pass kernels for log2(N) in {1 2 3 4 5 6 7 8}
#include "kernels_inst.cuh"
namespace nrb {
void register_col_a(PassTable &t)
{
    register_size<1, LAYOUT_COL>(t);
    register_size<2, LAYOUT_COL>(t);
    register_size<3, LAYOUT_COL>(t);
    register_size<4, LAYOUT_COL>(t);
    register_size<5, LAYOUT_COL>(t);
    register_size<6, LAYOUT_COL>(t);
    register_size<7, LAYOUT_COL>(t);
    register_size<8, LAYOUT_COL>(t);
}
} // namespace nrb
