// k_row_c.cu -- instantiates the LAYOUT_ROW FFT pass kernels for log2(N) in {12 13}
#include "kernels_inst.cuh"
namespace nrb {
void register_row_c(PassTable &t)
{
    register_size<12, LAYOUT_ROW>(t);
    register_size<13, LAYOUT_ROW>(t);
}
} // namespace nrb
