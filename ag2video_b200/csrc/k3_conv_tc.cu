// placeholder until the tcgen05 kernel lands (replaced in the next commit)
#include "k3_common.cuh"
namespace ag2v {
bool conv3x3_tc_supported(const ConvParams&, int) { return false; }
int conv3x3_tc(const ConvParams&, int, int, cudaStream_t) { return fail(AG2V_ERR_UNSUPPORTED, "tcgen05 conv not built"); }
}
