// Observation-window and reward assembly shared by all environment kernels.
//
// Restates (batched, one column = one actuator of one environment):
//   featurize       scripts/KS/setup/KSSetup.jl:190-229, KSglobalSetup.jl:211-249,
//                   scripts/Keller-Segel/setup/KellerSegelSetup.jl:265-316,
//                   scripts/Fluid/setup/FluidSetup.jl:204-245
//   reward_function KSSetup.jl:162-184, KSglobalSetup.jl:175-205,
//                   KellerSegelSetup.jl:241-263, FluidSetup.jl:188-202
// Indexing is integer-exact: Julia circshift(v, i)[j] == v[j - i] (periodic).
#pragma once
#include "common.cuh"

namespace pdeb200 {

__device__ __forceinline__ int wrap(int i, int n) { i %= n; return i < 0 ? i + n : i; }

// Sensor index feeding window row `row` (0-based, within one field) of actuator column c.
template <typename T>
__device__ __forceinline__ int window_sensor_index(const ObsRewardParams<T>& P, int row, int c) {
    const int m = P.a2s[c];
    const int h = P.window / 2;
    if (P.spa == 0) {                       // 1-D: rows i = -h..h, element = sens[m - i]
        return wrap(m - (row - h), P.n_sensors);
    }
    const int i = row / P.window - h;       // 2-D: i outer, j inner (FluidSetup.jl:220-222)
    const int j = row % P.window - h;
    const int a = m / P.spa, b = m % P.spa;
    return wrap(a - i, P.spa) * P.spa + wrap(b - j, P.spa);
}

// Writes the new state column and returns the per-actuator reward.
//   sens(field, i) -> raw <y_field, g_i> of this environment
//   a0 / da0: row 0 of action / delta_action for this column
//   action_col: global pointer to the column's a_rows action entries (memory rows)
//   state_col : global pointer to the column's obs_rows entries (read for temporal stacking)
//   fresh     : true for reset!/constructor semantics (the `isnothing(env)` branches)
template <typename T, typename SensFn>
__device__ __forceinline__ T assemble_column(const ObsRewardParams<T>& P, SensFn sens, int c, T a0, T da0,
                                             const T* action_col, T* state_col, bool fresh) {
    const int wrows = (P.spa == 0 ? P.window : P.window * P.window);
    const int block = wrows * P.fields;                 // rows produced by one featurize call
    if (P.temporal > 1 && !fresh) {
        // result = vcat(result, env.state[1:end-size(result)[1]-memory_size, :])
        for (int r = P.obs_rows - P.memory - 1; r >= block; --r) state_col[r] = state_col[r - block];
    }
    for (int f = 0; f < P.fields; ++f)
        for (int r = 0; r < wrows; ++r) {
            const T v = sens(f, window_sensor_index(P, r, c)) * P.obs_scale;
            state_col[f * wrows + r] = v;
            if (fresh) for (int k = 1; k < P.temporal; ++k) state_col[k * block + f * wrows + r] = v;
        }
    for (int k = 0; k < P.memory; ++k)
        state_col[P.obs_rows - P.memory + k] = fresh ? T(0) : action_col[P.a_rows - P.memory + k];
    const int m = P.a2s[c];
    const T raw = sens(0, m) - P.r_offset * P.sens_sum[m];
    T s;
    if (P.r_pow == T(2)) { const T g = P.r_gain * raw; s = g * g / P.r_div; }
    else s = pow_t<T>(fabs(P.r_gain * raw), P.r_pow) / P.r_div;
    return -fabs(s) - P.a_pun * a0 * a0 - P.da_pun * da0 * da0;
}

}  // namespace pdeb200
