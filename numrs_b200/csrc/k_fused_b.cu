// k_fused_b.cu -- fused z+y (rlft3) persistent kernels for nn3/2 = 256, nn2 in {256, 512, 1024}
#include "kernels_inst.cuh"
namespace nrb {
void register_fused_b()
{
    register_fused_zy<8, 8>();
    register_fused_zy<8, 9>();
    register_fused_zy<8, 10>();
}
} // namespace nrb
