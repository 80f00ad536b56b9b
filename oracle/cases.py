"""TEST INFRASTRUCTURE — the parity cases (small versions of BASELINE.json's configs plus
one case per loss / control / SDE / target combination the kernel claims to support).

Values follow the reference YAMLs cited per case (paths relative to /root/reference/conf).
`batch` is the golden-fixture batch; the -m gpu tests re-run the same case at larger B
against the oracle.
"""

LIN = lambda steps: {"steps": steps}  # noqa: E731

CASES = {
    # BASELINE cfg1: solver/basic_dis.yaml + target/dw_shift.yaml, loss.method=lv
    "dis_dw1_lv": dict(target="dw_shift", dim=1, sde="vp", prior="gauss", ctrl="lerp",
                       clip_model=1e4, clip_score=1e4, gate_bias=1.0, loss="time_reversal",
                       method="lv", max_rnd=None, timesteps=LIN(50), batch=64, seed=1),
    # BASELINE cfg2: basic_dis + GMM-40 "fab" d=2 (distr/gauss.py:42-47), lv
    "dis_gmm2_lv": dict(target="gmm40", dim=2, sde="vp", prior="gauss", ctrl="lerp",
                        clip_model=1e4, clip_score=1e4, gate_bias=1.0, loss="time_reversal",
                        method="lv", max_rnd=None, timesteps=LIN(100), batch=64, seed=1),
    # same, kl training semantics (rnd0 = 0, no Ito term): loss/time_reversal.yaml
    "dis_gmm2_kl": dict(target="gmm40", dim=2, sde="vp", prior="gauss", ctrl="lerp",
                        clip_model=1e4, clip_score=1e4, gate_bias=1.0, loss="time_reversal",
                        method="kl", max_rnd=None, timesteps=LIN(100), batch=64, seed=2),
    # BASELINE cfg3: solver/basic_pis.yaml + target/funnel.yaml (kl)
    "pis_funnel10_kl": dict(target="funnel", dim=10, sde="bm_pis", prior="delta", ctrl="score",
                            clip_model=1e4, clip_score=1e4, gate_bias=0.01, loss="reference_sde",
                            method="kl", max_rnd=None, timesteps=LIN(200), batch=64, seed=1),
    # BASELINE cfg4 / north-star headline: solver/dis.yaml (clips 10, max_rnd 1e8), GMM-40 d=50
    "dis_gmm50_lv": dict(target="gmm40", dim=50, sde="vp", prior="gauss", ctrl="lerp",
                         clip_model=10.0, clip_score=10.0, gate_bias=1.0, loss="time_reversal",
                         method="lv", max_rnd=1e8, timesteps=LIN(100), batch=48, seed=1),
    # solver/dds.yaml with the basic_dds grid (end 6.4 -> 129 steps; last dt ~ 0), funnel
    "dds_funnel10_lv": dict(target="funnel", dim=10, sde=None, prior="gauss", ctrl="score",
                            clip_model=10.0, clip_score=10.0, gate_bias=0.01, loss="exp_integrator",
                            method="lv", max_rnd=1e8, alpha=1.0, sigma=1.0,
                            timesteps=dict(rescale_t="cosine", end=6.4, dt=0.05), batch=64, seed=3),
    # solver/dds_euler.yaml: ReferenceSDELoss with reference_ctrl = sigma * prior score, VP
    "eulerdds_gmm2_lv": dict(target="gmm_rand", dim=2, sde="vp", prior="gauss", ctrl="score",
                             clip_model=10.0, clip_score=10.0, gate_bias=0.01, loss="reference_sde",
                             method="lv", euler_dds=True, max_rnd=1e8, timesteps=LIN(60), batch=64, seed=4),
    # solver/dis_no_score.yaml: ClippedCtrl; sde/const.yaml (ConstOU) to cover that coefficient family
    "dis_noscore_constou_gauss5": dict(target="gauss", dim=5, sde="const_ou", prior="gauss", ctrl="clipped",
                                       clip_model=10.0, loss="time_reversal", method="lv", max_rnd=1e8,
                                       timesteps=LIN(40), batch=64, seed=5),
    # model/lerp_prior.yaml, model/lerp_target.yaml; MultiWell; heterogeneous GMM; per-dim gate (lerp_dim)
    "dis_lerpprior_multiwell4": dict(target="multiwell", dim=4, sde="vp", prior="gauss", ctrl="lerp_prior",
                                     clip_model=1e4, clip_score=1e4, gate_bias=1.0, loss="time_reversal",
                                     method="kl_ito", max_rnd=None, timesteps=LIN(40), batch=64, seed=6),
    "dis_lerptarget_gmmrand3_dimgate": dict(target="gmm_rand", dim=3, sde="vp", prior="gauss", ctrl="lerp_target",
                                            clip_model=1e4, clip_score=5.0, gate_bias=1.0, gate_dim=3,
                                            loss="time_reversal", method="lv", max_rnd=None,
                                            clip_target=50.0, timesteps=LIN(40), batch=64, seed=7),
    # BASELINE cfg5 family: solver/dds.yaml (ScoreCtrl, clips 10, max_rnd 1e8) + target/nice.yaml, cosine grid.
    # NICE = 4 additive couplings (distr/nice.py); d > 64 and the coupling MLPs run on the wide (layered
    # tcgen05 GEMM) engine.  Small widths keep the fixtures small; the arithmetic path is the same.
    "dds_nice16_lv": dict(target="nice:24:3", dim=16, sde=None, prior="gauss", ctrl="score",
                          clip_model=10.0, clip_score=10.0, gate_bias=0.01, loss="exp_integrator",
                          method="lv", max_rnd=1e8, alpha=1.0, sigma=1.0,
                          timesteps=dict(rescale_t="cosine", end=3.2, dt=0.05), batch=48, seed=8),
    "dds_nice196_lv": dict(target="nice:72:5", dim=196, sde=None, prior="gauss", ctrl="score",
                           clip_model=10.0, clip_score=10.0, gate_bias=0.01, loss="exp_integrator",
                           method="lv", max_rnd=1e8, alpha=1.0, sigma=1.0,
                           timesteps=dict(rescale_t="cosine", end=1.6, dt=0.05), batch=24, seed=9),
    # a wide state with an analytic target: DIS+lv on a d=100 diagonal Gaussian (wide engine, Lerp control, VP)
    "dis_gauss100_lv": dict(target="gauss", dim=100, sde="vp", prior="gauss", ctrl="lerp",
                            clip_model=10.0, clip_score=10.0, gate_bias=1.0, loss="time_reversal",
                            method="lv", max_rnd=1e8, timesteps=LIN(30), batch=24, seed=10),
    # A mixture whose components differ in EVERY dimension (no shared-dimension factorisation): the DENSE instantiation of the
    # tensor-core rollout kernel (per-step score of all dims).  d=20 / 7 modes for the parity suite; d=50 / 40 modes is the
    # `gmm50dense` bench workload (bench.py) — the headline configuration without the zero-padded structure of GMM-40.
    "dis_gmmdense20_lv": dict(target="gmm_dense:7", dim=20, sde="vp", prior="gauss", ctrl="lerp",
                              clip_model=10.0, clip_score=10.0, gate_bias=1.0, loss="time_reversal",
                              method="lv", max_rnd=1e8, timesteps=LIN(40), batch=48, seed=16),
    "dis_gmmdense50_lv": dict(target="gmm_dense:40", dim=50, sde="vp", prior="gauss", ctrl="lerp",
                              clip_model=10.0, clip_score=10.0, gate_bias=1.0, loss="time_reversal",
                              method="lv", max_rnd=1e8, timesteps=LIN(100), batch=32, seed=17),
    "dds_gmmdense20_score": dict(target="gmm_dense:7", dim=20, sde=None, prior="gauss", ctrl="score",
                                 clip_model=10.0, clip_score=10.0, gate_bias=0.01, loss="exp_integrator",
                                 method="lv", max_rnd=1e8, alpha=1.0, sigma=1.0,
                                 timesteps=dict(rescale_t="cosine", end=3.2, dt=0.05), batch=48, seed=18),
    # ---- kl / kl_ito training gradients (backpropagation through time).  Together with dis_gmm2_kl,
    # pis_funnel10_kl and dis_lerpprior_multiwell4 above these cover every loss kind, the reference control of Euler-DDS,
    # active clips, and each target's second derivative (funnel, multiwell, Gauss; the GMM score is an autograd score
    # WITHOUT create_graph in the reference, distr/base.py:130-137, so it enters the adjoint as a constant).
    "dds_funnel10_kl": dict(target="funnel", dim=10, sde=None, prior="gauss", ctrl="score",
                            clip_model=10.0, clip_score=10.0, gate_bias=0.01, loss="exp_integrator",
                            method="kl", max_rnd=1e8, alpha=1.0, sigma=1.0,
                            timesteps=dict(rescale_t="cosine", end=3.2, dt=0.05), batch=48, seed=11),
    "eulerdds_gauss3_kl": dict(target="gauss", dim=3, sde="vp", prior="gauss", ctrl="score",
                               clip_model=10.0, clip_score=10.0, gate_bias=0.01, loss="reference_sde",
                               method="kl", euler_dds=True, max_rnd=1e8, timesteps=LIN(40), batch=48, seed=12),
    "dis_gmm50_kl": dict(target="gmm40", dim=50, sde="vp", prior="gauss", ctrl="lerp",
                         clip_model=10.0, clip_score=10.0, gate_bias=1.0, loss="time_reversal",
                         method="kl", max_rnd=1e8, timesteps=LIN(40), batch=32, seed=13),
    "dis_lerp_multiwell5_klito": dict(target="multiwell", dim=5, sde="vp", prior="gauss", ctrl="lerp",
                                      clip_model=2.0, clip_score=3.0, gate_bias=1.0, gate_dim=5, loss="time_reversal",
                                      method="kl_ito", max_rnd=None, clip_target=30.0, timesteps=LIN(40), batch=48, seed=14),
    # kl on the WIDE engine (NICE target / d > 64): the sweep's adjoint is a (B, d) fp32 plane, the dgrad chain per step GEMM launches
    "dds_nice16_kl": dict(target="nice:24:3", dim=16, sde=None, prior="gauss", ctrl="score",
                          clip_model=10.0, clip_score=10.0, gate_bias=0.01, loss="exp_integrator",
                          method="kl", max_rnd=1e8, alpha=1.0, sigma=1.0,
                          timesteps=dict(rescale_t="cosine", end=3.2, dt=0.05), batch=48, seed=19),
    "dis_gauss100_klito": dict(target="gauss", dim=100, sde="vp", prior="gauss", ctrl="lerp",
                               clip_model=10.0, clip_score=10.0, gate_bias=1.0, loss="time_reversal",
                               method="kl_ito", max_rnd=1e8, timesteps=LIN(30), batch=24, seed=20),
    # lv_traj (losses/oc.py:78-84): variance across the traj_per_sample trajectories of each initial point.  The fixture's x0
    # is the repeated batch (3 stacked copies of 24 points); the training call is given the 24 points.
    "dis_gmm2_lvtraj": dict(target="gmm40", dim=2, sde="vp", prior="gauss", ctrl="lerp",
                            clip_model=1e4, clip_score=1e4, gate_bias=1.0, loss="time_reversal",
                            method="lv_traj", traj_per_sample=3, max_rnd=None, timesteps=LIN(60), batch=72, seed=15),
}

# eval-mode variants: (case, compute_weights, return_traj)  — losses/oc.py:258-278, :371-392
EVAL_CASES = {
    "dis_gmm2_lv": [(True, True), (False, False)],
    "dis_dw1_lv": [(True, True)],
    "pis_funnel10_kl": [(True, False)],
    "dds_funnel10_lv": [(True, True)],
    "dds_nice16_lv": [(True, True)],
}

# Langevin dynamics (conf/solver/langevin.yaml: LangevinSDE diff_coeff 1, clip_score 1e5, EulerIntegrator dt 0.01) at
# small sizes; `eval_steps` output times — 7 is not a divisor of the step count, so outputs are interpolated
ULA_CASES = {
    "ula_gmm2": dict(target="gmm40", dim=2, terminal_t=2.0, dt=0.01, eval_steps=20, diff_coeff=1.0, clip_score=1e5, batch=64),
    "ula_funnel10": dict(target="funnel", dim=10, terminal_t=1.0, dt=0.01, eval_steps=7, diff_coeff=0.8, clip_score=3.0, batch=48),
    "ula_multiwell4": dict(target="multiwell", dim=4, terminal_t=0.5, dt=0.005, eval_steps=10, diff_coeff=1.0, clip_score=1e5, batch=32),
}

# EulerIntegrator.integrate on the OU family / ControlledSDE — the inference processes of TrainableDiff.compute_results
# (solver/oc.py:100-110: integrate(sde=inference_sde, ts=ts, x_init=target samples, timesteps=ts)) and a dt grid with
# interpolated outputs.  `grid`: "ts" = timesteps=ts as the solver passes them, or a dt for the integrator's own grid.
OU_CASES = {
    # DIS / Euler-DDS: inference_sde = VP(generative=False), no control (solver/oc.py:130-133, :287)
    "ou_vp_inference3": dict(sde="vp", generative=False, ctrl=None, dim=3, eval_steps=25, grid="ts", batch=64),
    # PIS: ControlledSDE(ScaledBM(generative=False), ctrl = PIS.inference_ctrl) (solver/oc.py:200-208) — a Brownian bridge to the origin
    "ou_pis_bridge4": dict(sde="bm_pis", generative=False, ctrl="pis", dim=4, eval_steps=40, grid="ts", batch=48),
    # generative ConstOU on the integrator's own dt grid: outputs interpolated (7 output intervals, 0.013-wide steps)
    "ou_const_dt2": dict(sde="const_ou", generative=True, ctrl=None, dim=2, eval_steps=7, grid=0.013, batch=32),
}

NOISE_SEED = 0x5DE5_0001
