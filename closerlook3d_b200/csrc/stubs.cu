// stubs.cu -- entry points declared in cl3d.h that are not implemented yet return CL3D_ERR_UNSUPPORTED.
#include "common.cuh"
using namespace cl3d;
#define STUB(name) { set_error(#name ": not implemented yet"); return CL3D_ERR_UNSUPPORTED; }
extern "C" size_t cl3d_grid_subsample_workspace_bytes(int, int, int) { return 256; }
extern "C" int cl3d_grid_subsample(const float*, const int*, int, int, int, float, float*, int*, void*, size_t, cl3d_stream_t) STUB(cl3d_grid_subsample)
extern "C" int cl3d_sgemm_nt(const float*, int, const float*, int, int, int, int, float*, int, cl3d_stream_t) STUB(cl3d_sgemm_nt)
extern "C" int cl3d_pwmlp_fwd_stats(const float*, const float*, const float*, const float*, const float*, const int*, int, int, int, int, int, float, float*, float*, unsigned char*, unsigned char*, float*, cl3d_stream_t) STUB(cl3d_pwmlp_fwd_stats)
extern "C" int cl3d_pwmlp_fwd_out(const float*, const float*, const float*, const float*, const float*, int, int, int, float*, cl3d_stream_t) STUB(cl3d_pwmlp_fwd_out)
extern "C" size_t cl3d_pwmlp_bwd_scratch_bytes(int, int, int, int, int) { return 256; }
extern "C" int cl3d_pwmlp_bwd(const float*, const float*, const float*, const float*, const float*, const float*, const float*, const int*, const float*, const float*, const unsigned char*, const unsigned char*, const float*, const float*, int, int, int, int, int, float, float*, size_t, float*, float*, float*, cl3d_stream_t) STUB(cl3d_pwmlp_bwd)
