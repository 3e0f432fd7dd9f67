// One (NT, WN) instantiation of the tensor-core round kernels; compiled once per pair with
// -DHMY_NT=<max n-tiles per warp> -DHMY_WN=<warps along the cluster axis> (harmonypy_b200/build.py).
#include "hmy_common.cuh"
#include "hmy_round.cuh"
#include "hmy_round_mma.cuh"
#include "hmy_ridge_mma.cuh"

#ifndef HMY_NT
#error "compile with -DHMY_NT=.. -DHMY_WN=.."
#endif
#define HMY_CATM2(a, b) hmy_bind_mma_##a##_##b
#define HMY_CATM(a, b) HMY_CATM2(a, b)

extern "C" void HMY_CATM(HMY_NT, HMY_WN)(const void** fns) {
    fns[0] = (const void*)k_round_mma<HMY_NT, HMY_WN, false>;
    fns[4] = (const void*)k_round_mma<HMY_NT, HMY_WN, true>;
    fns[1] = (const void*)k_round_mma_stage<HMY_NT, HMY_WN>;
    fns[2] = (const void*)k_ridge_moments_mma<HMY_NT, HMY_WN>;
    fns[3] = (const void*)k_ridge_apply_mma<HMY_NT, HMY_WN>;
}
