// mf_laplace.cu -- matrix-free Laplace/diffusion operator (placeholder until the cell kernel lands).
#include "mf.cuh"

using namespace mfmgb;

namespace mfmgb
{
int mf_apply(mfmgb_ctx *ctx, const mfmgb_mf *, const double *, Epi, const EpiArgs &)
{
  return fail(ctx, MFMGB_ERR_NOT_IMPLEMENTED, "matrix-free operator: not implemented yet");
}
} // namespace mfmgb

extern "C"
{
  MFMGB_API int mfmgb_mf_laplace_create(mfmgb_ctx *ctx, int, int, const int64_t *, const double *, const double *,
                                        const uint8_t *, mfmgb_mf **)
  {
    return fail(ctx, MFMGB_ERR_NOT_IMPLEMENTED, "mfmgb_mf_laplace_create: not implemented yet");
  }
  MFMGB_API int mfmgb_mf_destroy(mfmgb_ctx *, mfmgb_mf *) { return MFMGB_OK; }
  MFMGB_API int64_t mfmgb_mf_size(const mfmgb_mf *M) { return M ? M->n : 0; }
  MFMGB_API int mfmgb_mf_apply(mfmgb_ctx *ctx, const mfmgb_mf *, const double *, double *)
  {
    return fail(ctx, MFMGB_ERR_NOT_IMPLEMENTED, "mfmgb_mf_apply: not implemented yet");
  }
  MFMGB_API int mfmgb_mf_diagonal(mfmgb_ctx *ctx, const mfmgb_mf *, double *)
  {
    return fail(ctx, MFMGB_ERR_NOT_IMPLEMENTED, "mfmgb_mf_diagonal: not implemented yet");
  }
}
