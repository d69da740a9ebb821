// ggnn_tc.cu -- tcgen05 (BMP_MODE_BF16) GGNN encoder.  Placeholder until the
// tensor-core kernel lands: fails loudly, never falls back.
#include "common.cuh"
using namespace bmp;
int bmp_ggnn_forward_tc(const bmp_ggnn_fwd_t *a, void *stream) {
    (void)a; (void)stream;
    set_error("BMP_MODE_BF16 encoder is not built in this revision");
    return BMP_EINVAL;
}
