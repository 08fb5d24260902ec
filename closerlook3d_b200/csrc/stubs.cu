// stubs.cu -- entry points declared in cl3d.h that are not implemented yet return CL3D_ERR_UNSUPPORTED.
#include "common.cuh"
using namespace cl3d;
#define STUB(name) { set_error(#name ": not implemented yet"); return CL3D_ERR_UNSUPPORTED; }
extern "C" size_t cl3d_grid_subsample_workspace_bytes(int, int, int) { return 256; }
extern "C" int cl3d_grid_subsample(const float*, const int*, int, int, int, float, float*, int*, void*, size_t, cl3d_stream_t) STUB(cl3d_grid_subsample)
