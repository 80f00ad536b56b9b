/* sdes_b200.h — C ABI of the B200-native controlled-SDE rollout.
 *
 * This is the drop-in boundary for ONE path of juliusberner/sde_sampler: the T-step
 * rollout `simulate()` of the three optimal-control losses
 *     TimeReversalLoss.simulate               sde_sampler/losses/oc.py:156-230
 *     ReferenceSDELoss.simulate               sde_sampler/losses/oc.py:286-343
 *     ExponentialIntegratorSDELoss.simulate   sde_sampler/losses/oc.py:400-457
 * together with everything those loops call per step (control wrappers
 * models/reparam.py:13-200, FourierMLP/TimeEmbed models/mlp.py:43-122, SDE coefficients
 * eq/sdes.py:125-269, target / prior densities and scores distr/{gauss,double_well,
 * funnel}.py) and the reductions that follow them (BaseOCLoss.filter / compute_loss /
 * compute_results, losses/oc.py:50-123).
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; the binding a maintainer
 * adds is the ctypes stub in INTEGRATION.md (and sde_sampler_b200/_cabi.py is that stub).
 *
 * Conventions: plain pointers and sizes, no torch types. All array pointers are DEVICE
 * pointers (fp32 unless stated) on the device that is current when the call is made; all
 * work is enqueued on `stream` (a cudaStream_t passed as void*), nothing is allocated,
 * nothing synchronises the host.  State lives only in the descriptor: calls are re-entrant.
 * Return value 0 = ok, negative = error (sdes_last_error() gives the text, thread-local).
 */
#ifndef SDES_B200_H
#define SDES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDES_ABI_VERSION 3
#define SDES_CHANNELS 64     /* FourierMLP / TimeEmbed width (conf/model/base/fouriermlp.yaml:3) */
#define SDES_MAX_DIM 64      /* state dimension of the fused single-kernel engines (state in registers) */
#define SDES_MAX_WIDE_DIM 4096 /* state dimension of the wide (layered tcgen05 GEMM) engine: d > 64 or a NICE target */
#define SDES_MAX_HIDDEN 6    /* hidden layers per network                                    */
#define SDES_MAX_COMPONENTS 64 /* GMM components                                           */

/* which simulate() — losses/oc.py */
enum { SDES_LOSS_TIME_REVERSAL = 0, SDES_LOSS_REFERENCE_SDE = 1, SDES_LOSS_EXP_INTEGRATOR = 2 };
/* generative_ctrl wrapper — models/reparam.py */
enum { SDES_CTRL_CLIPPED = 0, SDES_CTRL_SCORE = 1, SDES_CTRL_LERP = 2, SDES_CTRL_LERP_PRIOR = 3,
       SDES_CTRL_LERP_TARGET = 4 };
/* sde coefficient family — eq/sdes.py */
enum { SDES_SDE_NONE = 0, SDES_SDE_VP = 1, SDES_SDE_CONST_OU = 2 /* ConstOU and ScaledBM */ };
/* analytic target — distr/*.py.  Gauss / IsotropicGauss targets are passed as GMM with K=1. */
enum { SDES_TARGET_GMM = 0, SDES_TARGET_MULTIWELL = 1 /* DoubleWell = n_dw=d=1 */, SDES_TARGET_FUNNEL = 2,
       SDES_TARGET_NICE = 3 /* distr/nice.py:127-263, wide engine */ };

/* flags */
#define SDES_F_RND0_ZERO      (1u << 0) /* rnd starts at 0 instead of log p_prior(x0) (oc.py:168-172)   */
#define SDES_F_COMPUTE_ITO    (1u << 1) /* accumulate the Ito integral (oc.py:218-219,:330-331,:440-443) */
#define SDES_F_SUB_DIV_INT    (1u << 2) /* eval: rnd -= int_s^t div(mu)  (oc.py:210-211)               */
#define SDES_F_RETURN_TRAJ    (1u << 3) /* write xs (T+1,B,d)  (oc.py:221-222,:228-229)                  */
#define SDES_F_NOISE_FROM_HBM (1u << 4) /* parity mode: eps read from `noise` (T,B,d) instead of Philox */
#define SDES_F_REFERENCE_CTRL (1u << 5) /* Euler-DDS reference_ctrl = sigma * prior score (solver/oc.py:305-306) */
#define SDES_F_HAS_GATE       (1u << 6) /* ctrl.score_model (a TimeEmbed gate) is present                */
#define SDES_F_MLP_SIMT       (1u << 7) /* evaluate the control MLP with fp32 FFMA instead of tcgen05    */
#define SDES_F_TRAJ_TILED     (1u << 8) /* with RETURN_TRAJ (d <= SDES_MAX_DIM): xs is written row-tiled,
                                           [T+1][ceil(B/128)][jdim][128] floats with jdim = 8/16/32/48/56/64 >= d
                                           (coalesced for thread-per-trajectory kernels); the layout
                                           sdes_rollout_lv_grad reads when the same flag is set on its descriptor */

#define SDES_F_KEEP_FOR_GRAD  (1u << 9) /* wide engine: keep what sdes_rollout_lv_grad needs INSIDE the workspace (the
                                           operand image of the state at every step, the per-step gate cotangent
                                           sums); the gradient call must then be given the same workspace */

#define SDES_F_KEEP_SCORE     (1u << 10) /* with KEEP_FOR_GRAD, for sdes_rollout_kl_grad on the wide engine: also keep the target
                                           score of every step and of the terminal state (T + 1 planes of (B, d) fp32) */

/* Layout of the flat fp32 parameter blob `params` (torch (out,in) row-major weights, C = 64):
 *   FourierMLP (models/mlp.py:85-122)
 *     in_w[C*d] in_b[C]
 *     te_phase[C]  te_h0_w[C*2C] te_h0_b[C]  { te_hk_w[C*C] te_hk_b[C] } x (te_hidden-1)  te_out_w[C*C] te_out_b[C]
 *     { h_w[C*C] h_b[C] } x n_hidden
 *     out_w[d*C] out_b[d]
 *   gate TimeEmbed (models/mlp.py:43-82), only with SDES_F_HAS_GATE
 *     g_phase[C]  g_h0_w[C*2C] g_h0_b[C]  { g_hk_w[C*C] g_hk_b[C] } x (gate_hidden-1)  g_out_w[gate_dim*C] g_out_b[gate_dim]
 */
typedef struct SdesRolloutDesc {
    uint32_t struct_bytes;   /* = sizeof(SdesRolloutDesc), checked */
    uint32_t abi_version;    /* = SDES_ABI_VERSION, checked */
    int32_t loss_kind, ctrl_kind, sde_kind, target_kind;
    uint32_t flags;
    int32_t dim;             /* d: 1..SDES_MAX_DIM on the fused engines, up to SDES_MAX_WIDE_DIM on the wide engine */
    int32_t n_steps;         /* T >= 1; ts has T+1 entries */
    int32_t n_hidden;        /* FourierMLP hidden (C->C) layers = num_layers-2 */
    int32_t te_hidden;       /* hidden layers of FourierMLP.timestep_embed (>=1; reference: 1) */
    int32_t gate_hidden;     /* hidden layers of the gate TimeEmbed (>=1; reference: 3) */
    int32_t gate_dim;        /* 1 or d */
    int32_t n_components;    /* GMM: K */
    int32_t n_double_wells;  /* MULTIWELL */
    int64_t batch;           /* B trajectories in this call (this rank's shard) */
    uint64_t traj_offset;    /* global index of trajectory 0 (Philox counter; shard-invariant noise) */
    uint64_t seed;           /* Philox key */
    /* +/-inf = no clip (utils/common.py:83-84) */
    float clip_model, clip_score, clip_target, scale_score;
    float alpha, sigma;      /* ExponentialIntegratorSDELoss (oc.py:395-398) */
    /* eq/sdes.py: VP uses beta_min/beta_max/scale_diff/terminal_t/sign; CONST_OU drift_coeff/diff_coeff/sign */
    float beta_min, beta_max, scale_diff, terminal_t, sde_sign, drift_coeff, diff_coeff;
    /* targets: MULTIWELL separation/shift; FUNNEL variance; all: log_norm_const */
    float separation, shift, variance, log_norm_const;
    const float* ts;         /* (T+1) */
    const float* params;     /* blob, layout above */
    int64_t n_params;        /* its length in floats, checked against the layout */
    const float* gmm_loc;    /* (K,d) */
    const float* gmm_scale;  /* (K,d) */
    const float* gmm_weights;/* (K) unnormalised mixture weights, or NULL = uniform */
    const float* prior_loc;  /* (d) diagonal Gaussian prior (initial cost, Lerp* ctrl, reference ctrl), may be NULL if unused */
    const float* prior_scale;
    const float* ref_loc;    /* (d) reference_distr of REFERENCE_SDE / EXP_INTEGRATOR terminal cost */
    const float* ref_scale;
    const float* x0;         /* (B,d) row-major */
    const float* noise;      /* (T,B,d) with SDES_F_NOISE_FROM_HBM, else NULL */
    float* x_T;              /* (B,d) out */
    float* rnd;              /* (B)   out */
    float* xs;               /* (T+1,B,d) out with SDES_F_RETURN_TRAJ, else NULL */
    void* workspace;         /* >= sdes_workspace_bytes(desc), 256-byte aligned */
    size_t workspace_bytes;
    /* NICE target (target_kind = SDES_TARGET_NICE; distr/nice.py): `nice_couplings` additive couplings, each a
     * ReLU MLP  half -> mid -> ... -> mid -> half  with `nice_hidden` hidden layers (nice_hidden + 1 Linear),
     * half = d/2; coupling c transforms the even units if (nice_mask_config + c) % 2 == 1, else the odd units
     * (Coupling.forward :64-95); then z = h * exp(scale) with a standard-logistic latent (:21-29, :109-124).
     * nice_params (fp32, torch (out,in) row-major):
     *   { in_w[mid*half] in_b[mid]  { mid_w[mid*mid] mid_b[mid] } x (nice_hidden-1)  out_w[half*mid] out_b[half] } x couplings
     *   scale[d]                                                                                                  */
    int32_t nice_couplings, nice_mid, nice_hidden, nice_mask_config;
    const float* nice_params;
    int64_t n_nice_params;
    /* Optional output of the tensor-core fused engine (d <= SDES_MAX_DIM, no SDES_F_MLP_SIMT), NULL = not wanted:
     * gate_cot (T, B) = ito coefficient x sum_j eps_j x [scale_score (sigma) clip(inner_j)], the ungated score part of the
     * control paired with the step's noise — what d rnd_b / d gate(s) is for the log-variance losses.  A training forward
     * that keeps it spares sdes_rollout_lv_grad the re-evaluation of the target score (SdesLvGradDesc.gate_cot). */
    float* gate_cot;
    /* Optional output of the same engine for the kl / kl_ito gradient, NULL = not wanted: score_keep, in the layout of xs
     * (steps 0 .. T-1; honours SDES_F_TRAJ_TILED), holds the UNGATED score part of the control at every (step, trajectory,
     * dimension): scale_score (sigma) clip(inner_j, clip_score).  With it sdes_rollout_kl_grad runs its whole reverse sweep
     * as one persistent kernel (SdesLvGradDesc.score_keep) whenever the score term's x-derivative is local per dimension. */
    float* score_keep;
} SdesRolloutDesc;

/* ABI version of the loaded library (== SDES_ABI_VERSION of the header it was built from). */
int sdes_version(void);

/* Text of the last error on this thread ("" if none). */
const char* sdes_last_error(void);

/* Bytes of device scratch one sdes_rollout_fwd call needs (per-step tables, re-laid-out weights). */
size_t sdes_workspace_bytes(const SdesRolloutDesc* desc);

/* The whole rollout: replaces `loss.simulate(ts, x, ...)` (losses/oc.py:156,:286,:400).
 * d <= SDES_MAX_DIM with an analytic target: enqueues (1) a small prologue kernel that hoists everything
 * x-independent into per-step tables and (2) ONE persistent kernel that carries each trajectory through all
 * T steps.  d > SDES_MAX_DIM or a NICE target (BASELINE cfg5): the wide engine — the state lives in HBM and
 * every Linear of the control MLP and of the NICE couplings (forward and the input-gradient backward that
 * gives the target score) is a tcgen05 GEMM launch with a fused epilogue, one fused update kernel per step. */
int sdes_rollout_fwd(const SdesRolloutDesc* desc, void* stream);

/* 1 if the tcgen05 (tensor-core) engine handles this descriptor, 0 if only the fp32-FFMA engine
 * (SDES_F_MLP_SIMT) does.  sdes_rollout_fwd never switches engines by itself: asking for the
 * tensor-core engine on an unsupported descriptor is an error. */
int sdes_tcgen05_supported(const SdesRolloutDesc* desc);

/* Gradient of the log-variance loss with respect to the control network — `loss.backward()` of
 * `Trainable.step` (solver/base.py:404-407) for loss.method = lv.  In the lv losses the state is
 * driven by the detached control (losses/oc.py:60-64) and the running cost has zero derivative, so
 *     d loss / d theta = sum_b w[b] sum_s J_theta g(s, x_{b,s})^T c_{b,s},   c = eps sqrt(dt)  (DDS: sigma beta_k eps)
 * — one backward pass of the control MLP over all B*T rows of the stored trajectory, no backpropagation through
 * time.  `desc` is the descriptor of the forward call (same seed / traj_offset / noise so that eps is re-drawn
 * identically; x0/x_T/rnd/xs of desc are ignored; desc->workspace must hold sdes_lv_grad_workspace_bytes).
 * Fused engines: `xs` is the stored trajectory; wide engine (d > SDES_MAX_DIM or NICE): the forward ran with
 * SDES_F_KEEP_FOR_GRAD and desc->workspace is that same workspace.  Outputs (overwritten):
 *   grad_params (n_params floats, the layout of `params`): the gradient of EVERY control parameter — the x-dependent
 *               layers (in_w, h_w/h_b, out_w/out_b) from the B*T-row passes, in_b and the two TimeEmbed networks
 *               (timestep embedding, gate) from the per-step cotangents below
 *   grad_emb    (T, 64)        d loss / d (timestep_embed(s_i) + in_b)        (intermediate, also returned)
 *   grad_gate   (T, gate_dim)  d loss / d gate(s_i)                           (intermediate, also returned) */
#define SDES_GRAD_TARGET_SCORE_CONST (1u << 0) /* kl gradient: the target score inside the control is a constant of the
                                                  graph — the reference's autograd score without create_graph
                                                  (Distribution.score distr/base.py:130-137 called from
                                                  models/reparam.py:60,:135): GMM targets */
#define SDES_GRAD_SCORE_DETACHED     (1u << 1) /* kl gradient: ctrl.detach_score=True — the whole score part (target and
                                                  prior score) is evaluated on x.detach() (models/reparam.py:58,:134) */
#define SDES_GRAD_LAYERWISE_SWEEP     (1u << 2) /* tcgen05 engine, lv and kl gradients: do NOT take the one-kernel path
                                                  (csrc/sdes_grad_fused.cuh) — run the layer-by-layer GEMM passes, and for
                                                  kl the step-by-step sweep with one GEMM launch per transposed layer —
                                                  same result; cross-check and A/B measurements */
typedef struct SdesLvGradDesc {
    uint32_t struct_bytes;   /* = sizeof(SdesLvGradDesc), checked */
    uint32_t flags;          /* SDES_GRAD_* (the score-term flags: sdes_rollout_kl_grad only) */
    const float* xs;         /* (T+1, B, d) trajectory written by the forward call with SDES_F_RETURN_TRAJ */
    const float* w;          /* (B) d loss / d rnd_b (0 for filtered trajectories) */
    float* grad_params;
    float* grad_emb;
    float* grad_gate;        /* NULL when the control has no gate */
    int64_t chunk_rows;      /* rows (trajectory, step) per pass; 0 = default (2^20) */
    const float* gate_cot;   /* lv, scalar gate, fused engines: (T, B) from the forward's SdesRolloutDesc.gate_cot, or NULL
                                (the gate gradient is then recomputed from the target score) */
    const float* score_keep; /* kl / kl_ito, tensor-core fused engine: the forward's SdesRolloutDesc.score_keep, or NULL
                                (the sweep then runs step by step and re-evaluates the target score) */
} SdesLvGradDesc;

size_t sdes_lv_grad_workspace_bytes(const SdesRolloutDesc* desc, const SdesLvGradDesc* g);
int sdes_rollout_lv_grad(const SdesRolloutDesc* desc, const SdesLvGradDesc* g, void* stream);

/* Gradient of the kl / kl_ito losses with respect to the control network — `loss.backward()` of `Trainable.step`
 * (solver/base.py:404-407) for loss.method = kl | kl_ito.  Here the state is driven by the control WITH
 * its graph (`sde_ctrl = generative_ctrl`, losses/oc.py:180,:305,:421): backpropagation through time, done as a
 * discrete adjoint over the stored trajectory.  tcgen05 engine: the sweep runs inside the gradient's chunk loop (chunks
 * in reverse time order) — per step one elementwise kernel (control cotangent, score-term x-derivatives) and the dgrad
 * GEMM chain on that step's row tiles, accumulating into the fp32 adjoint; weight gradients once per chunk.  With
 * SDES_F_MLP_SIMT: one thread-per-trajectory reverse-sweep kernel (fp32 FFMA) writes the cotangent of the control at
 * every (trajectory, step) into the workspace and the batched pass of sdes_rollout_lv_grad consumes it.  Same descriptors and outputs as sdes_rollout_lv_grad; `w` = d loss / d rnd_b
 * (sdes_kl_weights); g->flags = SDES_GRAD_*.  A GMM target with more than one component requires
 * SDES_GRAD_TARGET_SCORE_CONST (the reference's semantics) or SDES_GRAD_SCORE_DETACHED.
 * Wide engine (d > SDES_MAX_DIM or a NICE target): the forward ran with SDES_F_KEEP_FOR_GRAD | SDES_F_KEEP_SCORE in the same
 * workspace; the sweep is one elementwise kernel + the dgrad GEMM chain (incl. W_in^T, accumulating into the fp32 adjoint)
 * per step inside the reverse chunk loop.  A NICE / multi-component GMM score is a constant of the graph (the reference's
 * autograd score without create_graph); a single Gaussian target is differentiated analytically; wide funnel / multi-well
 * targets are refused. */
size_t sdes_kl_grad_workspace_bytes(const SdesRolloutDesc* desc, const SdesLvGradDesc* g);
int sdes_rollout_kl_grad(const SdesRolloutDesc* desc, const SdesLvGradDesc* g, void* stream);

/* EulerIntegrator.integrate (eq/integrator.py:79-127) for a LangevinSDE (eq/sdes.py:38-65) — the unadjusted
 * Langevin sampler of LangevinSolver.run (solver/langevin.py:34-63), SURVEY §8f-3:
 *   x <- x + clip(score(x) diff_coeff^2 / 2, clip_score) (t - s) + diff_coeff eps sqrt(t - s)   over `timesteps`,
 * output at the times `out_ts` by linear interpolation inside the step that covers them (interpolate(), :66-77).
 * The target is described by the rollout descriptor's target fields (target_kind, gmm_*, separation, ...), plus dim,
 * batch, seed, traj_offset, flags & SDES_F_NOISE_FROM_HBM with noise (n_steps, B, d), workspace (>=
 * sdes_integrate_workspace_bytes).  d <= SDES_MAX_DIM, analytic targets.  One launch for the whole chain. */
typedef struct SdesIntegrateDesc {
    uint32_t struct_bytes;   /* = sizeof(SdesIntegrateDesc), checked */
    int32_t n_steps;         /* integration steps; timesteps has n_steps + 1 entries */
    int32_t n_out;           /* output times */
    float diff_coeff, clip_score /* +inf = none */, eps /* EulerIntegrator.eps, 1e-8 */;
    const float* timesteps;
    const float* out_ts;
    const float* x_init;     /* (B, d) */
    float* xs_out;           /* (n_out, B, d) */
    int32_t noise_is_increment; /* with SDES_F_NOISE_FROM_HBM: 0 = `noise` holds standard normals (scaled by sqrt(t - s) here),
                                   1 = it holds the Brownian increments themselves — the `bm(s, t)` values of
                                   eq/integrator.py:116-119, evaluated by the caller */
    int32_t reserved;
} SdesIntegrateDesc;

size_t sdes_integrate_workspace_bytes(const SdesRolloutDesc* target_desc);
int sdes_langevin_integrate(const SdesRolloutDesc* target_desc, const SdesIntegrateDesc* g, void* stream);

/* EulerIntegrator.integrate (eq/integrator.py:79-127) for the OU family (eq/sdes.py:66-269: VP, ConstOU, ScaledBM, generative
 * or not) and for a ControlledSDE (eq/sdes.py:272-305) whose control is the score of a Gaussian marginal — the inference
 * processes of TrainableDiff.compute_results (solver/oc.py:100-110, called with timesteps = ts; PIS: :204-208), SURVEY §8f-3:
 *   x <- x + (mu_i x + sigma_i c_i(x)) (t - s) + sigma_i dW_i,
 *   c_i(x)_j = csig_i min((cloc[i, j] - x_j) cinv_i, cmax)          (absent when cloc is NULL)
 * with the x-independent per-step coefficients in `tab` (n_steps x 8 floats: mu, sigma, csig, cinv, 4 reserved) — the
 * caller evaluates the reference's own coefficient functions on the grid once.  dW_i = eps sqrt(t - s) with eps from the
 * Philox stream (counter layout of the rollout), or from `noise` (n_steps, B, d) as standard normals / increments.
 * Output at `out_ts` by the reference's linear interpolation.  Any d; one launch for the whole chain. */
typedef struct SdesAffineIntegrateDesc {
    uint32_t struct_bytes;   /* = sizeof(SdesAffineIntegrateDesc), checked */
    int32_t dim;
    int64_t batch;
    int32_t n_steps, n_out;
    float eps;               /* EulerIntegrator.eps */
    float cmax;              /* upper clip of the control's score (PIS: 1e5); +inf = none */
    const float* timesteps;  /* (n_steps + 1) */
    const float* out_ts;     /* (n_out) */
    const float* tab;        /* (n_steps, 8) */
    const float* cloc;       /* (n_steps, d) or NULL */
    const float* x_init;     /* (B, d) */
    const float* noise;      /* (n_steps, B, d) or NULL (in-kernel Philox) */
    int32_t noise_is_increment;
    int32_t reserved;
    uint64_t seed, traj_offset;
    float* xs_out;           /* (n_out, B, d) */
} SdesAffineIntegrateDesc;

int sdes_affine_integrate(const SdesAffineIntegrateDesc* g, void* stream);

/* The expectation estimates of LangevinSolver.run (solver/langevin.py:50-54; EXPECTATION_FNS, distr/base.py:12-17) over the
 * rows of `xs` (n_rows, d) — the caller passes xs[burn_steps:] flattened: out4 (device, 4 doubles) = means over rows of
 * sum_j x^2, sum_j |x|, sum_j x, sum_j (x^2 - x).  One launch. */
int sdes_expectations(const float* xs, int64_t n_rows, int32_t dim, double* out4, void* stream);

/* Statistics of rnd that BaseOCLoss.filter/compute_loss/compute_results reduce to
 * (losses/oc.py:50-123).  out_stats (device, 8 doubles):
 *   [0] n_kept  [1] sum(rnd | kept)  [2] sum(rnd^2 | kept)  [3] max(-rnd | kept)
 *   [4] sum(exp(-rnd - [3]) | kept)  [5] n_total
 *   [6] unbiased variance of the kept rnd = the lv loss, [7] their mean = the kl loss — of THIS call's shard (a
 *       multi-rank caller recomputes both from the combined [0..2])
 * keep-mask by mask_mode: 0 = isfinite(rnd), 1 = rnd < max_rnd (oc.py:50-58), 2 = keep all
 * (compute_results applies no mask, oc.py:94-123); `sample_mask` (B bytes, 0 = drop; may be
 * NULL) is the result of the caller's filter_samples(x_T) (oc.py:53-55), AND-ed in.
 * Ranks combine these without touching rnd again: n, sums add; max is max; the exp-sums are
 * rescaled by exp([3]_rank - [3]_global) and added. */
int sdes_rnd_stats(const float* rnd, int64_t batch, int mask_mode, float max_rnd,
                   const uint8_t* sample_mask, double* out_stats, void* stream);

/* The rank-combining step of the statistics above as ONE launch: `gathered` (device, world x 8 doubles, the output of an
 * all-gather of every rank's out_stats) -> out_stats of the global batch (all eight slots, [6] / [7] recomputed from the
 * combined [0..2]).  Ranks that kept nothing contribute max = -inf and an exp-sum of 0; a NaN max propagates. */
int sdes_merge_stats(const double* gathered, int32_t world, double* out_stats, void* stream);

/* Importance weights exp(-rnd - max(-rnd)) (losses/oc.py:104-105); the shift is stats[3], read on device. */
int sdes_weights(const float* rnd, int64_t batch, const double* stats, float* weights, void* stream);

/* lv_traj (losses/oc.py:78-84): rnd holds traj_per_sample trajectories for each of n_samples initial points, laid out
 * (traj_per_sample, n_samples) as `x.repeat(traj_per_sample, 1, 1)` produces (:240-241).  A sample is kept when every one
 * of its trajectories passes the mask; out3 (device, 3 doubles) = [sum over kept samples of the unbiased variance across
 * their trajectories, kept samples, all samples].  loss = [0] / [1]; ranks add the three numbers. */
int sdes_lv_traj_stats(const float* rnd, int64_t n_samples, int32_t traj_per_sample, int mask_mode, float max_rnd,
                       const uint8_t* sample_mask, double* out3, void* stream);

/* Cotangent of the log-variance loss with respect to rnd: w[b] = upstream * 2 (rnd_b - mean) / (n - 1) for kept b (the
 * mask of sdes_rnd_stats), 0 otherwise; `stats` are the (rank-combined) statistics, `upstream` a device scalar
 * (d objective / d loss, e.g. scale_loss; NULL = 1).  Input `w` of sdes_rollout_lv_grad. */
int sdes_lv_weights(const float* rnd, int64_t batch, int mask_mode, float max_rnd, const uint8_t* sample_mask,
                    const double* stats, const float* upstream, float* w, void* stream);

/* Cotangent of the lv_traj loss (losses/oc.py:78-84) with respect to rnd, layout as sdes_lv_traj_stats:
 * w[t, i] = upstream * 2 (rnd[t, i] - mean_t rnd[., i]) / (traj_per_sample - 1) / kept_samples for kept samples i, else 0;
 * `out3` are the (rank-combined) numbers of sdes_lv_traj_stats.  Input `w` of sdes_rollout_lv_grad. */
int sdes_lv_traj_weights(const float* rnd, int64_t n_samples, int32_t traj_per_sample, int mask_mode, float max_rnd,
                         const uint8_t* sample_mask, const double* out3, const float* upstream, float* w, void* stream);

/* Cotangent of the kl loss (mean of the kept rnd, losses/oc.py:90) with respect to rnd: w[b] = upstream / n_kept for kept b,
 * 0 otherwise.  Input `w` of sdes_rollout_kl_grad. */
int sdes_kl_weights(const float* rnd, int64_t batch, int mask_mode, float max_rnd, const uint8_t* sample_mask,
                    const double* stats, const float* upstream, float* w, void* stream);

/* ---- the caller's side of the rollout ------------------------------------------------------------ */

/* x0 ~ prior: IsotropicGauss.sample (distr/gauss.py:228-242) — `loc + scale * randn` or, with truncate_quartile,
 * nn.init.trunc_normal_(mean, std, a, b) (uniform on [2 Phi(a')-1, 2 Phi(b')-1], erfinv, scale, shift, clamp) — drawn from
 * the Philox stream keyed by `seed` with counter (traj_offset + b, 0xFFFFFFFF, dim/4), so shards of a batch draw what the
 * whole batch would.  `uniforms` (B,d in [0,1), may be NULL) replaces the Philox uniforms of the truncated path (parity
 * hook).  out (B,d). */
int sdes_sample_gauss_prior(float* out, int64_t batch, int32_t dim, float mean, float std, int32_t truncated, float a, float b,
                            uint64_t seed, uint64_t traj_offset, const float* uniforms, void* stream);

/* The tail of Trainable.step (solver/base.py:409-439) on flat fp32 buffers of n parameters, without a host sync:
 *   loss_ok = isfinite(loss) or |loss| <= max_loss;  grad_ok = all-finite or inf-norm <= max_grad      (:410-421)
 *   if both: clip_grad_norm_(max_norm = grad_clip_norm, L2; conf/utils/grad_clip.yaml), torch.optim.Adam step
 *   (lr, betas, eps, L2 weight_decay; conf/solver/oc_base.yaml:26-29), EMA.update (solver/base.py:620-684: copy until
 *   update_after_step, then every update_every calls shadow -= (1 - decay_t)(shadow - param) with the warm-up decay of
 *   get_current_decay);  else: skipped-step counter += 1.
 * +inf for max_loss / max_grad / grad_clip_norm = not set; ema_shadow NULL = no EMA; loss NULL = no loss check.
 * state: 8 doubles on the device, zero-initialised by the caller, carried from call to call:
 *   [0] optimizer steps taken  [1] skipped steps  [2] EMA num_updates  [3] gradient L2 norm of this call
 *   [4] gradient inf-norm  [5] 1 if this call stepped  [6] EMA decay used (-1: no EMA update)  [7] clip coefficient
 * The learning-rate / parameter schedulers stay with the caller (they are Python objects); `lr` is read per call. */
typedef struct SdesTrainerStepDesc {
    uint32_t struct_bytes;   /* = sizeof(SdesTrainerStepDesc), checked */
    uint32_t reserved;
    int64_t n;
    float* params;
    const float* grads;
    float* exp_avg;
    float* exp_avg_sq;
    float* ema_shadow;
    const float* loss;
    float lr, beta1, beta2, eps, weight_decay;
    float max_loss, max_grad, grad_clip_norm;
    double ema_decay, ema_inv_gamma, ema_power, ema_min_value;   /* doubles: the reference evaluates the decay with Python floats */
    int32_t ema_update_after_step, ema_update_every;
    double* state;
    void* workspace;         /* >= sdes_trainer_workspace_bytes() */
    size_t workspace_bytes;
} SdesTrainerStepDesc;
size_t sdes_trainer_workspace_bytes(void);
int sdes_trainer_step(const SdesTrainerStepDesc* desc, void* stream);

/* Sample statistics of get_metrics (eval/metrics.py:120-131) in one pass: out (device, 4 + 2 d doubles) =
 *   [0] sum w  [1] sum w^2  [2] B  [3] 0  [4 + j] sum_b x_bj  [4 + d + j] sum_b x_bj^2
 * (ESS = [0]^2 / [1]; stddev_j, avg_j from the moments).  weights may be NULL. */
int sdes_eval_moments(const float* samples, const float* weights, int64_t batch, int32_t dim, double* out, void* stream);

/* The noise stream on its own: eps (T,B,d) exactly as the fused kernel draws it in registers
 * (Philox4x32-10 keyed by seed, counter (traj_offset+b, step, dim/4) + Box-Muller).  Test hook. */
int sdes_philox_normal(uint64_t seed, uint64_t traj_offset, int64_t batch, int32_t n_steps,
                       int32_t dim, float* out, void* stream);

/* Tensor-core self test: D[128,N] = A[128,K] * W[N,K]^T through the rollout's own tcgen05 path
 * (A split into bf16 hi / lo and staged into TMEM with tcgen05.st, W hi / lo images in shared memory,
 * three kind::f16 passes per k-step, tcgen05.ld).  K multiple of 8 in [8,64], N multiple of 16 in
 * [16,64].  mode must be 0 (the one split the rollout uses).  Test hook. */
int sdes_tcgen05_selftest(const float* a, const float* w, float* d, int32_t k, int32_t n, int32_t mode, void* stream);

/* y[i] = GELU(x[i]) as the layered GEMM epilogues evaluate it (x Phi(x), erfc by Abramowitz-Stegun 7.1.26
 * on MUFU.RCP / MUFU.EX2).  Test hook for the accuracy claim in DESIGN.md. */
int sdes_gelu_probe(const float* x, float* y, int64_t n, void* stream);

/* The same for the packed two-lane GELU of the persistent rollout kernel's epilogue (logistic form
 * x / (1 + 2^(-x P(x^2))), all FP32 work as f32x2 instructions).  n must be even. */
int sdes_gelu_pair_probe(const float* x, float* y, int64_t n, void* stream);

/* Kernel launches performed by this process through the library since load (for bench accounting). */
int64_t sdes_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SDES_B200_H */
