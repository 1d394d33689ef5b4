// k_fused_a.cu -- fused z+y (rlft3) persistent kernels for nn3/2 = 128, nn2 in {256, 512, 1024}
#include "kernels_inst.cuh"
namespace nrb {
void register_fused_a()
{
    register_fused_zy<7, 8>();
    register_fused_zy<7, 9>();
    register_fused_zy<7, 10>();
}
} // namespace nrb
