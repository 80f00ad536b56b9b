// sdes_integrate.cu — EulerIntegrator.integrate (eq/integrator.py:79-127) for LangevinSDE (eq/sdes.py:38-65):
// the unadjusted Langevin sampler of `LangevinSolver.run` (solver/langevin.py:34-63), SURVEY §8f-3.
//
//     for (s, t) in zip(timesteps[:-1], timesteps[1:]):
//         xt = xs + clip(score(xs) * sigma^2 / 2, clip_score) * (t - s) + sigma * randn * sqrt(t - s)
//         every output time tau in ts with tau <= t + eps that has not been written yet gets
//         lerp(xs, xt, (tau - s) / (t - s))                                   (interpolate(), :66-77)
//
// One thread per trajectory, the state in registers, the analytic target score evaluated in place
// (sdes_step.cuh), Philox noise drawn per step with the rollout's counter layout — the whole chain of
// (typically 10 000) steps is one launch; the reference launches ~10 kernels per step.
#include "sdes_step.cuh"

namespace sdes {


template <int DPAD>
__global__ void __launch_bounds__(128) langevin_kernel(const __grid_constant__ IntegrateParams a) {
    extern __shared__ __align__(16) float smem_f[];
    const SdesRolloutDesc& d = a.d;
    const float* ws = reinterpret_cast<const float*>(d.workspace);
    const int dim = d.dim, K = d.n_components, tid = threadIdx.x;
    const int K2 = (K + 1) & ~1;
    float* s_mu = smem_f;
    float* s_h = s_mu + K2 * DPAD;
    float* s_c = s_h + K2 * DPAD;
    for (int e = tid; e < K2 * DPAD; e += blockDim.x) {
        s_mu[e] = ws[a.ws.gmm_mu + e];
        s_h[e] = ws[a.ws.gmm_h + e];
    }
    for (int e = tid; e < 64; e += blockDim.x) s_c[e] = ws[a.ws.gmm_c + e];
    __syncthreads();
    TargetSmem tsm{s_mu, s_h, s_c, reinterpret_cast<const uint32_t*>(ws + a.ws.counter)[1], s_c, s_c};

    const int64_t B = d.batch, row = (int64_t)blockIdx.x * blockDim.x + tid;
    const bool valid = row < B;
    const int64_t rrow = valid ? row : B - 1;
    float x[DPAD];
#pragma unroll
    for (int j = 0; j < DPAD; ++j) x[j] = (j < dim) ? __ldg(a.x_init + rrow * dim + j) : 0.f;
    const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)rrow);
    const uint32_t k0 = (uint32_t)d.seed, k1 = (uint32_t)(d.seed >> 32);
    const bool from_hbm = (d.flags & SDES_F_NOISE_FROM_HBM) != 0;
    const float half_s2 = a.diff_coeff * a.diff_coeff / 2.0f;
    int cnt = 0;
    for (int i = 0; i < a.n_steps; ++i) {
        const float s = a.timesteps[i], t = a.timesteps[i + 1];
        const float dt = t - s, sq = sqrtf(dt);
        float sc[DPAD], xn[DPAD];
        target_eval<DPAD, true>(d, x, sc, tsm);
#pragma unroll
        for (int q = 0; q < DPAD / 4; ++q) {
            float e[4] = {0.f, 0.f, 0.f, 0.f};
            if (4 * q < dim) {
                if (from_hbm) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) e[r] = (4 * q + r < dim) ? d.noise[((int64_t)i * B + rrow) * dim + 4 * q + r] : 0.f;
                } else {
                    const float4 n4 = normal4_call(k0, k1, traj, (uint32_t)i, (uint32_t)q);
                    e[0] = n4.x; e[1] = n4.y; e[2] = n4.z; e[3] = n4.w;
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = 4 * q + r;
                // drift = clip(score * sigma^2 / 2) (eq/sdes.py:54-61); noise = randn * sqrt(t - s) (integrator.py:116)
                xn[j] = (j < dim) ? x[j] + clipf(sc[j] * half_s2, a.clip_score) * dt + a.diff_coeff * (e[r] * sq) : 0.f;
            }
        }
        // output times falling into (.., t + eps]: linear interpolation between xs and xt (integrator.py:66-77, :121-123)
        while (cnt < a.n_out && a.out_ts[cnt] <= t + a.eps) {
            const float wgt = (a.out_ts[cnt] - s) / dt;
            if (valid) {
                float* o = a.xs_out + ((int64_t)cnt * B + row) * dim;
#pragma unroll
                for (int j = 0; j < DPAD; ++j)
                    if (j < dim) o[j] = torch_lerp(x[j], xn[j], wgt);
            }
            ++cnt;
        }
#pragma unroll
        for (int j = 0; j < DPAD; ++j) x[j] = xn[j];
    }
}

template <int DPAD>
static cudaError_t launch_langevin_t(const IntegrateParams& a, cudaStream_t stream) {
    const int K2 = (a.d.n_components + 1) & ~1;
    const size_t smem = (2 * (size_t)K2 * DPAD + 64) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(langevin_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    langevin_kernel<DPAD><<<(int)((a.d.batch + 127) / 128), 128, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_langevin(const IntegrateParams& a, cudaStream_t stream) {
    switch (a.ws.dpad) {
        case 4: return launch_langevin_t<4>(a, stream);
        case 8: return launch_langevin_t<8>(a, stream);
        case 12: return launch_langevin_t<12>(a, stream);
        case 16: return launch_langevin_t<16>(a, stream);
        case 32: return launch_langevin_t<32>(a, stream);
        case 52: return launch_langevin_t<52>(a, stream);
        case 64: return launch_langevin_t<64>(a, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sdes
