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
        const float dt = t - s, nsc = (from_hbm && a.noise_is_increment) ? 1.0f : sqrtf(dt);
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
                xn[j] = (j < dim) ? x[j] + clipf(sc[j] * half_s2, a.clip_score) * dt + a.diff_coeff * (e[r] * nsc) : 0.f;
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

// ---------------------------------------------------------------------------- affine SDEs
// EulerIntegrator.integrate for the OU family and Gaussian-score ControlledSDEs (include/sdes_b200.h: sdes_affine_integrate).
// The drift is elementwise in x, so a thread owns four dimensions of one trajectory for the whole chain (the Philox counter
// layout of the rollout: (trajectory, step, dim / 4)); adjacent threads own adjacent dimension chunks of a row, so the
// interpolated outputs are written as contiguous 16-byte pieces of the (n_out, B, d) tensor.
__global__ void __launch_bounds__(256) affine_integrate_kernel(const SdesAffineIntegrateDesc g) {
    const int nchunk = (g.dim + 3) / 4;
    const int64_t total = g.batch * nchunk;
    const uint32_t k0 = (uint32_t)g.seed, k1 = (uint32_t)(g.seed >> 32);
    const bool vec = (g.dim & 3) == 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / nchunk;
        const int q = (int)(e - row * nchunk), j0 = 4 * q;
        float x[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) x[r] = (j0 + r < g.dim) ? __ldg(g.x_init + row * g.dim + j0 + r) : 0.f;
        const uint32_t traj = (uint32_t)(g.traj_offset + (uint64_t)row);
        int cnt = 0;
        for (int i = 0; i < g.n_steps; ++i) {
            const float s = g.timesteps[i], t = g.timesteps[i + 1], dt = t - s;
            const float* tab = g.tab + (int64_t)i * 8;
            const float mu = tab[0], sigma = tab[1], csig = tab[2], cinv = tab[3];
            float n[4];
            if (g.noise != nullptr) {
                const float nsc = g.noise_is_increment ? 1.0f : sqrtf(dt);
#pragma unroll
                for (int r = 0; r < 4; ++r) n[r] = (j0 + r < g.dim) ? g.noise[((int64_t)i * g.batch + row) * g.dim + j0 + r] * nsc : 0.f;
            } else {
                const float4 n4 = normal4_call(k0, k1, traj, (uint32_t)i, (uint32_t)q);
                const float sq = sqrtf(dt);
                n[0] = n4.x * sq; n[1] = n4.y * sq; n[2] = n4.z * sq; n[3] = n4.w * sq;  // torch.randn * sqrt(t - s)  (integrator.py:116)
            }
            float xn[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float drift = mu * x[r];  // OU.drift (eq/sdes.py:94-95)
                if (g.cloc != nullptr && j0 + r < g.dim) {
                    // ControlledSDE.f_and_g (eq/sdes.py:296-305): + sde_diff * ctrl, ctrl = diff * score.clip(max=cmax) (solver/oc.py:206-208)
                    const float score = (g.cloc[(int64_t)i * g.dim + j0 + r] - x[r]) * cinv;
                    drift += sigma * (csig * fminf(score, g.cmax));
                }
                xn[r] = x[r] + drift * dt + sigma * n[r];
            }
            while (cnt < g.n_out && g.out_ts[cnt] <= t + g.eps) {  // interpolate(), integrator.py:66-77, :121-123
                const float wgt = (g.out_ts[cnt] - s) / dt;
                float* o = g.xs_out + ((int64_t)cnt * g.batch + row) * g.dim + j0;
                if (vec) {
                    *reinterpret_cast<float4*>(o) = make_float4(torch_lerp(x[0], xn[0], wgt), torch_lerp(x[1], xn[1], wgt),
                                                                 torch_lerp(x[2], xn[2], wgt), torch_lerp(x[3], xn[3], wgt));
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (j0 + r < g.dim) o[r] = torch_lerp(x[r], xn[r], wgt);
                }
                ++cnt;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) x[r] = xn[r];
        }
    }
}

cudaError_t launch_affine_integrate(const SdesAffineIntegrateDesc& g, cudaStream_t stream) {
    const int64_t total = g.batch * ((g.dim + 3) / 4);
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    affine_integrate_kernel<<<(int)blocks, 256, 0, stream>>>(g);
    return cudaGetLastError();
}

// Means over rows of sum_j x^2, sum_j |x|, sum_j x (EXPECTATION_FNS, distr/base.py:12-17): per-thread fp64 partials over a
// grid-stride sweep of the flat tensor, one atomicAdd per block and quantity.
__global__ void __launch_bounds__(256) expectations_kernel(const float* __restrict__ xs, int64_t n, double inv_rows, double* __restrict__ out) {
    __shared__ double s_p[3][8];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)xs[i];
        a += v * v;
        b += fabs(v);
        c += v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) { s_p[0][threadIdx.x >> 5] = a; s_p[1][threadIdx.x >> 5] = b; s_p[2][threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = c = 0.0;
        for (int w = 0; w < 8; ++w) { a += s_p[0][w]; b += s_p[1][w]; c += s_p[2][w]; }
        atomicAdd(out + 0, a * inv_rows);
        atomicAdd(out + 1, b * inv_rows);
        atomicAdd(out + 2, c * inv_rows);
        atomicAdd(out + 3, (a - c) * inv_rows);
    }
}

cudaError_t launch_expectations(const float* xs, int64_t n_rows, int dim, double* out4, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(out4, 0, 4 * sizeof(double), stream);
    if (e != cudaSuccess || n_rows == 0) return e;
    const int64_t n = n_rows * dim;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    expectations_kernel<<<(int)blocks, 256, 0, stream>>>(xs, n, 1.0 / (double)n_rows, out4);
    return cudaGetLastError();
}

}  // namespace sdes
