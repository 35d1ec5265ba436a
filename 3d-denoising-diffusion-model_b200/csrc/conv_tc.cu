// placeholder until the tcgen05 kernel lands (next commit)
#include "kernels.h"
namespace ddpm3d {
bool conv_tc_eligible(const ConvArgs&) { return false; }
int conv_tc(const ConvArgs&, cudaStream_t) { set_error("conv_tc: not built"); return DDPM3D_ERR_ARG; }
}  // namespace ddpm3d
