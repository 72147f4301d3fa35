// kernels_kg.cu -- instantiates every kernel for ONE lane-group width (compile with -DQS_KG=1|2|4|8|16|32).
#include "launch.h"

#ifndef QS_KG
#error "compile with -DQS_KG=<lanes per env>"
#endif

namespace qs {
namespace {

constexpr int KG = QS_KG;

template <typename... Args>
void set_attr(size_t bytes, void (*kernel)(Args...))
{
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    // the kernels stream their state (no L1 reuse): give the whole unified L1/shared array to shared memory so that the
    // observation tiles never limit the number of resident blocks
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// feature sets of the upstream step kernel: bit 0 obstacles, bit 1 downwash, bit 2 formation scenarios (never with obstacles)
#define QS_FEAT_SWITCH(FEATv, CALL)                       \
    switch (FEATv) {                                      \
        case 0: { constexpr int FEAT = 0; CALL; } break;  \
        case 1: { constexpr int FEAT = 1; CALL; } break;  \
        case 2: { constexpr int FEAT = 2; CALL; } break;  \
        case 3: { constexpr int FEAT = 3; CALL; } break;  \
        case 4: { constexpr int FEAT = 4; CALL; } break;  \
        default: { constexpr int FEAT = 6; CALL; } break; \
    }
// the persistent form exists for the feature sets without formation scenarios
#define QS_FEAT_SWITCH_PERSIST(FEATv, CALL)               \
    switch (FEATv) {                                      \
        case 0: { constexpr int FEAT = 0; CALL; } break;  \
        case 1: { constexpr int FEAT = 1; CALL; } break;  \
        case 2: { constexpr int FEAT = 2; CALL; } break;  \
        default: { constexpr int FEAT = 3; CALL; } break; \
    }

void prepare(int feat, size_t smem_plain, size_t smem_persist, int block, int *persist_blocks_per_sm, bool fork)
{
    *persist_blocks_per_sm = 0;
    if (fork) {
        set_attr(smem_plain, fork_step_kernel<KG>);
        set_attr(smem_plain, fork_reset_kernel<KG>);
        return;
    }
    QS_FEAT_SWITCH(feat, set_attr(smem_plain, step_kernel<KG, false, FEAT>));
    set_attr(smem_plain, reset_kernel<KG, false, false>);
    set_attr(smem_plain, reset_kernel<KG, true, false>);
    set_attr(smem_plain, reset_kernel<KG, false, true>);
    if (feat < 4 && smem_persist > 0 && smem_persist <= 227 * 1024) {
        QS_FEAT_SWITCH_PERSIST(feat, set_attr(smem_persist, step_kernel<KG, true, FEAT>);
                       cudaOccupancyMaxActiveBlocksPerMultiprocessor(persist_blocks_per_sm, step_kernel<KG, true, FEAT>, block, smem_persist));
    }
}

void step(bool persist, int feat, LaunchShape s, cudaStream_t st, const DevConst &c, const DevPtrs &P, const float4 *actions, float *obs,
          float *rew, uint8_t *done, float *term_obs, uint8_t *reset_success)
{
    if (persist && feat < 4) {
        QS_FEAT_SWITCH_PERSIST(feat, (step_kernel<KG, true, FEAT><<<s.grid, s.block, s.smem, st>>>(c, P, actions, obs, rew, done, term_obs, reset_success)));
    } else {
        QS_FEAT_SWITCH(feat, (step_kernel<KG, false, FEAT><<<s.grid, s.block, s.smem, st>>>(c, P, actions, obs, rew, done, term_obs, reset_success)));
    }
}

void reset(int feat, LaunchShape s, cudaStream_t st, const DevConst &c, const DevPtrs &P, const uint8_t *mask, float *obs)
{
    if (feat & 1) reset_kernel<KG, true, false><<<s.grid, s.block, s.smem, st>>>(c, P, mask, obs);
    else if (feat & 4) reset_kernel<KG, false, true><<<s.grid, s.block, s.smem, st>>>(c, P, mask, obs);
    else reset_kernel<KG, false, false><<<s.grid, s.block, s.smem, st>>>(c, P, mask, obs);
}

void fork_step(LaunchShape s, cudaStream_t st, const DevConst &c, const ForkConst &f, const DevPtrs &P, const ForkPtrs &F,
               const float2 *actions, float *obs, float *rew, uint8_t *done, float *term_obs, uint8_t *reset_success)
{
    fork_step_kernel<KG><<<s.grid, s.block, s.smem, st>>>(c, f, P, F, actions, obs, rew, done, term_obs, reset_success);
}

void fork_reset(LaunchShape s, cudaStream_t st, const DevConst &c, const ForkConst &f, const DevPtrs &P, const ForkPtrs &F,
                const uint8_t *mask, float *obs)
{
    fork_reset_kernel<KG><<<s.grid, s.block, s.smem, st>>>(c, f, P, F, mask, obs);
}

const KgLaunchers table = { prepare, step, reset, fork_step, fork_reset };

}  // namespace

#define QS_CAT2(a, b) a##b
#define QS_CAT(a, b) QS_CAT2(a, b)
const KgLaunchers &QS_CAT(launchers_kg, QS_KG)() { return table; }

}  // namespace qs
