// One NC instantiation (NC = 16-cluster column chunks of the scoring tile) of the tcgen05 / tensor-memory
// round kernel; compiled once per NC with -DHMY_TC5_NC=<4|7|8> (harmonypy_b200/build.py).
#include "hmy_common.cuh"
#include "hmy_round.cuh"
#include "hmy_round_mma.cuh"
#include "hmy_round_tc5.cuh"

#ifndef HMY_TC5_NC
#error "compile with -DHMY_TC5_NC=.."
#endif
#define HMY_CATT2(a) hmy_bind_tc5_##a
#define HMY_CATT(a) HMY_CATT2(a)

extern "C" void HMY_CATT(HMY_TC5_NC)(const void** fns) {
    fns[0] = (const void*)k_round_tc5<HMY_TC5_NC, false>;
    fns[1] = (const void*)k_round_tc5<HMY_TC5_NC, true>;
}
