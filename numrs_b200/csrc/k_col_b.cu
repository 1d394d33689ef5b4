// k_col_b.cu -- instantiates the LAYOUT_COL FFT pass kernels for log2(N) in {9 10}
#include "kernels_inst.cuh"
namespace nrb {
void register_col_b(PassTable &t)
{
    register_size<9, LAYOUT_COL>(t);
    register_size<10, LAYOUT_COL>(t);
}
} // namespace nrb
