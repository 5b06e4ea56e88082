// placeholder: replaced by the tcgen05 kernel (work in progress)
#include "common.cuh"
#include "garment4d_b200.h"
G4D_API int g4d_sa_mlp_k0(int c_in) { return ((9 + c_in) + 15) / 16 * 16; }
G4D_API size_t g4d_sa_mlp_param_bytes(const g4d_sa_mlp_desc* d) { (void)d; return 0; }
G4D_API int g4d_sa_mlp_pack_params(const g4d_sa_mlp_desc*, const float*, const float*, const float*, const float*, const float*, const float*, void*) { return g4d::bad_arg("sa_mlp: not built yet"); }
G4D_API int g4d_sa_mlp_max(const g4d_sa_mlp_desc*, const void*, int, int, int, const float*, const float*, const int*, const void*, float*, void*, int, int, void*) { return g4d::bad_arg("sa_mlp: not built yet"); }
