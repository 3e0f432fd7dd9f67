// One (KPT, JPW) instantiation of the streaming kernels; compiled once per pair with
// -DHMY_KPT=<clusters per lane> -DHMY_JPW=<PCs per warp> (see harmonypy_b200/build.py).
#include "hmy_common.cuh"
#include "hmy_round.cuh"
#include "hmy_ridge.cuh"

#ifndef HMY_KPT
#error "compile with -DHMY_KPT=.. -DHMY_JPW=.."
#endif
#define HMY_CAT2(a, b, c) hmy_bind_##a##_##b
#define HMY_CAT(a, b) HMY_CAT2(a, b, 0)

extern "C" void HMY_CAT(HMY_KPT, HMY_JPW)(const void** fns) {
    fns[0] = (const void*)k_round<HMY_KPT, HMY_JPW>;
    fns[1] = (const void*)k_round_stage<HMY_KPT, HMY_JPW>;
    fns[2] = (const void*)k_ridge_moments<HMY_KPT, HMY_JPW>;
    fns[3] = (const void*)k_ridge_apply<HMY_KPT, HMY_JPW>;
}
