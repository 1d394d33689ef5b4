// k_fused_c.cu -- fused z+y (rlft3) persistent kernels for nn3/2 = 512, nn2 in {256, 512, 1024}
#include "kernels_inst.cuh"
namespace nrb {
void register_fused_c()
{
    register_fused_zy<9, 8>();
    register_fused_zy<9, 9>();
    register_fused_zy<9, 10>();
}
} // namespace nrb
