// k_row_a.cu -- instantiates the LAYOUT_ROW FFT pass kernels for log2(N) in {1 2 3 4 5 6 7 8}
#include "kernels_inst.cuh"
namespace nrb {
void register_row_a(PassTable &t)
{
    register_size<1, LAYOUT_ROW>(t);
    register_size<2, LAYOUT_ROW>(t);
    register_size<3, LAYOUT_ROW>(t);
    register_size<4, LAYOUT_ROW>(t);
    register_size<5, LAYOUT_ROW>(t);
    register_size<6, LAYOUT_ROW>(t);
    register_size<7, LAYOUT_ROW>(t);
    register_size<8, LAYOUT_ROW>(t);
}
} // namespace nrb
