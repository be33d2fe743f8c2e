// uvs_marg.cu — construction of the next marginalization prior (placeholder until the device path lands).
#include "uvs_handle.h"
#include "uvs_kernels.h"

int uvs_marginalize_impl(UvsHandle *h, int window_index, int flag, UvsPrior *out) {
  (void)window_index; (void)flag; (void)out;
  return uvs::handle_fail(h, UVS_ERR_UNSUPPORTED, "uvs_marginalize: not implemented yet");
}
