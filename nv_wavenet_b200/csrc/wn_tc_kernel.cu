// wn_tc_kernel.cu -- fp16 tensor-core (tcgen05 / TMEM / TMA) kernel family.  (stub until implemented)
#include "wn_common.h"

bool wn_tc_supported(int, int, int, int, int) { return false; }
size_t wn_tc_image_bytes(int, int, int, int) { return 256; }
cudaError_t wn_tc_pack(void*, const WnParams&, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t wn_launch_tc(const WnParams&, const void*, cudaStream_t, WnLaunchInfo*) { return cudaErrorNotSupported; }
