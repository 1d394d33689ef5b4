// k_col_c.cu -- instantiates the LAYOUT_COL FFT pass kernels for log2(N) in {11 12}
#include "kernels_inst.cuh"
namespace nrb {
void register_col_c(PassTable &t)
{
    register_size<11, LAYOUT_COL>(t);
    register_size<12, LAYOUT_COL>(t);
}
} // namespace nrb
