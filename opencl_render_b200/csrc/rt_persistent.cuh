// rt_persistent.cuh -- persistent-thread kernel (placeholder: falls back to nothing yet)
#pragma once
#include <string>
#include "rt_kernels.cuh"
namespace oclr {
inline bool launch_persistent(const SceneView& S, const FrameView& F, int smCount, uint32_t* workCounter, Counters* dcnt,
                              cudaStream_t st, uint32_t& launches, std::string& err) {
    (void)smCount; (void)workCounter;
    const size_t shBytes = sizeof(float) * 3 * (S.n + 1);
    dim3 grid((F.cam.width + 15) / 16, (launch_rows(F) + 7) / 8);
    if (dcnt) raytrace_simple_kernel<true><<<grid, 128, shBytes, st>>>(S, F, dcnt);
    else raytrace_simple_kernel<false><<<grid, 128, shBytes, st>>>(S, F, dcnt);
    launches = 1;
    (void)err;
    return true;
}
}
