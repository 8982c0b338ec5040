// skyjo_step_inst.cu -- instantiates the fused step / observe kernels for one player count.
// Compiled once per N with -DSKYJO_N=<1..12> (build.py) so the 12 translation units build
// in parallel.
#include "skyjo_step.cuh"

#ifndef SKYJO_N
#error "compile with -DSKYJO_N=<num_players>"
#endif

#define SKYJO_CAT2(a, b) a##b
#define SKYJO_CAT(a, b) SKYJO_CAT2(a, b)

namespace skyjo {

cudaError_t SKYJO_CAT(launch_step_, SKYJO_N)(const StepParams &p, bool indirect, bool policy, cudaStream_t s) {
    constexpr int N = SKYJO_N;
    const dim3 grid((unsigned)(p.tiles > 0 ? p.tiles : p.Bpad / TILE)), block(TILE);
    const size_t smem = (size_t)TILE * ((indirect ? 31 : 19 + 12 * N) + 26);
    // programmatic stream serialization: the grid may be scheduled while its predecessor drains
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = p.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (indirect) {
        if (policy) return cudaLaunchKernelEx(&cfg, step_kernel<N, true, true>, p);
        return cudaLaunchKernelEx(&cfg, step_kernel<N, true, false>, p);
    }
    if (policy) return cudaLaunchKernelEx(&cfg, step_kernel<N, false, true>, p);
    return cudaLaunchKernelEx(&cfg, step_kernel<N, false, false>, p);
}

cudaError_t SKYJO_CAT(launch_rollout_, SKYJO_N)(const StepParams &p, const RolloutParams &r, bool indirect,
                                                 cudaStream_t s) {
    constexpr int N = SKYJO_N;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.Bpad / TILE));
    cfg.blockDim = dim3(TILE);
    cfg.dynamicSmemBytes = (size_t)TILE * ((indirect ? 31 : 19 + 12 * N) + 26);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = p.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (indirect) return cudaLaunchKernelEx(&cfg, rollout_kernel<N, true>, p, r);
    return cudaLaunchKernelEx(&cfg, rollout_kernel<N, false>, p, r);
}

cudaError_t SKYJO_CAT(launch_observe_, SKYJO_N)(const StepParams &p, bool indirect, int agent, int8_t *obs,
                                                 int8_t *mask, int reset_outputs, int bulk_ok, cudaStream_t s) {
    constexpr int N = SKYJO_N;
    const dim3 grid((unsigned)(p.Bpad / TILE)), block(TILE);
    const size_t smem = (size_t)TILE * ((indirect ? 31 : 19 + 12 * N) + 26);
    if (indirect) observe_kernel<N, true><<<grid, block, smem, s>>>(p, agent, obs, mask, reset_outputs, bulk_ok);
    else observe_kernel<N, false><<<grid, block, smem, s>>>(p, agent, obs, mask, reset_outputs, bulk_ok);
    return cudaGetLastError();
}

}  // namespace skyjo
