#include "conv_tc.h"
namespace cindm {
int launch_conv_tc(const ConvTcLaunch&, cudaStream_t) { return fail(-99, "tcgen05 conv engine not built yet"); }
}
