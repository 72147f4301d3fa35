// launch.h -- per-lane-group launchers.  The kernels are templates on KG (lanes per env) and, for the upstream step kernel,
// on the persistent form and the compile-time feature set; every KG is compiled in its own translation unit
// (kernels_kg.cu with -DQS_KG=n) so that the library builds in parallel.  quadsim.cu only sees these plain functions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "quadsim_kernels.cuh"
#include "fork_kernels.cuh"

namespace qs {

struct LaunchShape { int grid, block; size_t smem; };

// one table entry per KG in {1, 2, 4, 8, 16, 32}
struct KgLaunchers {
    // set shared-memory attributes; for the persistent form return the resident blocks per SM (0 if it does not fit)
    void (*prepare)(int feat, size_t smem_plain, size_t smem_persist, int block, int *persist_blocks_per_sm, bool fork);
    void (*step)(bool persist, int feat, LaunchShape s, cudaStream_t st, const DevConst &c, const DevPtrs &P, const float4 *actions,
                 float *obs, float *rew, uint8_t *done, float *term_obs, uint8_t *reset_success);
    void (*reset)(int feat, LaunchShape s, cudaStream_t st, const DevConst &c, const DevPtrs &P, const uint8_t *mask, float *obs);
    void (*fork_step)(LaunchShape s, cudaStream_t st, const DevConst &c, const ForkConst &f, const DevPtrs &P, const ForkPtrs &F,
                      const float2 *actions, float *obs, float *rew, uint8_t *done, float *term_obs, uint8_t *reset_success);
    void (*fork_reset)(LaunchShape s, cudaStream_t st, const DevConst &c, const ForkConst &f, const DevPtrs &P, const ForkPtrs &F,
                       const uint8_t *mask, float *obs);
};

const KgLaunchers &launchers_kg1();
const KgLaunchers &launchers_kg2();
const KgLaunchers &launchers_kg4();
const KgLaunchers &launchers_kg8();
const KgLaunchers &launchers_kg16();
const KgLaunchers &launchers_kg32();

}  // namespace qs
