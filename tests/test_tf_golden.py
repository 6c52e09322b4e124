"""Pin BOTH the CPU oracle (oracle/tf_half.py) and the CUDA path to golden outputs of the reference's UNMODIFIED TF1
graph code (tests/golden/tf_half_graph.npz, written by oracle/make_golden.py::gen_tf_half_graph, which runs
meta_algos/{base,pro_mp,trpo_maml,vpg_maml}.py, optimizers/*, policies/* from /root/reference on the torch-backed
`tensorflow` stand-in).  Inputs are regenerated from the seeds in oracle/tf_cases.py.

Each quantity exists twice in the fixture: 'f32' (the reference's dtype) and 'f64' (the same graph without rounding).
The bar (BASELINE.json north_star): float32 results within 1e-4 relative of the reference - checked against the f64
values, with the f32 reference's own distance to f64 as the noise scale where float32 itself is the limit
(finite-difference Hx).  Includes the BASELINE.json full-size shapes: configs[1] 40 x 2000 (Do 2 / Da 2),
configs[2] 40 x 4000 (Do 17 / Da 6), configs[3] TRPO-MAML 40 x 2000."""
import os

import numpy as np
import pytest

from oracle import tf_cases

GOLD = None


def _gold(golden_dir):
    global GOLD
    if GOLD is None:
        GOLD = np.load(os.path.join(golden_dir, 'tf_half_graph.npz'))
    return GOLD


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def close(a, b, rtol, atol=0.0):
    return abs(float(a) - float(b)) <= atol + rtol * abs(float(b))


ALL = list(tf_cases.CASES)
PROMP = [n for n in ALL if tf_cases.CASES[n]['algo'] == 'promp']
TRPO = [n for n in ALL if tf_cases.CASES[n]['algo'] == 'trpo']
VPG = [n for n in ALL if tf_cases.CASES[n]['algo'] == 'vpg']


# ================================================================================================ CPU: oracle vs reference graph
def _oracle_data(case, dt):
    import torch
    N = case['N']
    return [dict(obs=torch.tensor(p['obs'], dtype=dt), act=torch.tensor(p['act'], dtype=dt),
                 adv=torch.tensor(p['adv'], dtype=dt), mean=torch.tensor(p['mean'], dtype=dt),
                 log_std=torch.tensor(p['log_std'], dtype=dt)[:, None, :].expand(-1, N, -1),
                 adj_avg_rewards=torch.tensor(p['adj_avg_rewards'], dtype=dt)) for p in case['phases']]


@pytest.mark.parametrize('name', ALL)
@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_oracle_matches_reference_graph(golden_dir, name, tag):
    """oracle/tf_half.py (the restatement bench.py's CPU legs and the other GPU tests use) == the unmodified reference
    graph: adapted parameters, objective, KLs, second-order meta-gradient."""
    import torch
    from oracle import tf_half as th
    G = _gold(golden_dir)
    case = tf_cases.make_case(name)
    if case['M'] >= 40 and tag == 'f32':
        pytest.skip("full-size cases: the float64 evaluation is the pin (keeps the CPU suite short)")
    dt = torch.float32 if tag == 'f32' else torch.float64
    tol = 2e-5 if tag == 'f32' else 1e-6           # f64: the reference's inner_lr / log(2 pi) are float32-rounded constants
    pre = '%s/%s/' % (name, tag)
    dims = (case['Do'], case['Da'], (case['hidden'],) * 2)
    data = _oracle_data(case, dt)
    inner = case.get('inner_type', 'likelihood_ratio')
    theta = torch.tensor(case['theta'], dtype=dt)
    keep = G[name + '/keep_tasks']
    cur = theta[None].expand(case['M'], -1).contiguous()
    for s in range(case['S'] - 1):
        cur = th.adapt(cur, data[s], dims, 0.1, inner)
        want = G[pre + 'adapt%d_tasks' % s]
        assert rel_err(cur.numpy()[keep] - case['theta'], want - case['theta']) < tol
        delta = cur.numpy().astype(np.float64) - case['theta'].astype(np.float64)
        np.testing.assert_allclose(np.sqrt((delta ** 2).sum(1)), G[pre + 'adapt%d_delta_norm' % s], rtol=10 * tol)
    t = theta.clone().requires_grad_(True)
    obj, ikl, okl = th.meta_objective(t, data, dims, 0.1, case['algo'], 0.3, [5e-4] * (case['S'] - 1), inner,
                                      exploration=case.get('exploration', False))
    (g,) = torch.autograd.grad(obj, t)
    assert close(obj.detach(), G[pre + 'loss'], tol, 1e-7)
    assert rel_err(g.numpy(), G[pre + 'grad']) < tol
    if case['algo'] == 'promp':
        np.testing.assert_allclose(ikl.detach().numpy(), G[pre + 'inner_kl'], rtol=10 * tol, atol=1e-9)
    if case['algo'] in ('promp', 'trpo'):
        assert close(okl.detach(), G[pre + 'outer_kl'], 10 * tol, 1e-9)


@pytest.mark.parametrize('name', [n for n in PROMP if tf_cases.CASES[n]['M'] < 40])
def test_oracle_adam_trajectory_matches_reference_graph(golden_dir, name):
    """5 epochs of the reference's MAMLPPOOptimizer (unmodified; TF1 Adam rule from the stand-in) vs oracle promp_optimize."""
    import torch
    from oracle import tf_half as th
    G = _gold(golden_dir)
    case = tf_cases.make_case(name)
    dims = (case['Do'], case['Da'], (case['hidden'],) * 2)
    for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
        pre = '%s/%s/' % (name, tag)
        data = _oracle_data(case, dt)
        theta0 = torch.tensor(case['theta'], dtype=dt)
        for K in (1, 5):
            adam = th.TF1Adam(theta0.numel(), dtype=dt)
            got, st = th.promp_optimize(theta0.clone(), data, dims, adam, 0.1, 0.3, [5e-4] * (case['S'] - 1), K)
            want = G[pre + 'theta_after_adam%d' % K]
            upd_err = rel_err(got.numpy() - case['theta'], want - case['theta'])
            assert upd_err < (5e-3 if tag == 'f32' else 1e-5), (tag, K, upd_err)
            if K == 5:
                assert close(st['loss_after'], G[pre + 'loss_after'], 1e-4, 1e-6)
                np.testing.assert_allclose(st['inner_kls'], G[pre + 'inner_kl_after'], rtol=1e-3, atol=1e-8)


@pytest.mark.parametrize('name', [n for n in TRPO if tf_cases.CASES[n]['M'] < 40])
def test_oracle_trpo_matches_reference_graph(golden_dir, name):
    """TRPOMAML + ConjugateGradientOptimizer + FiniteDifferenceHvp (unmodified reference) vs the oracle, in float64
    where the finite difference is not rounding noise: constraint gradient, Hx, CG direction, accepted parameters."""
    import torch
    from oracle import tf_half as th
    G = _gold(golden_dir)
    case = tf_cases.make_case(name)
    if case.get('exploration'):
        pytest.skip("TRPOMAMLOracle has no exploration term (the E-MAML objective/gradient is covered above)")
    dims = (case['Do'], case['Da'], (case['hidden'],) * 2)
    pre = name + '/f64/'
    data = _oracle_data(case, torch.float64)
    o = th.TRPOMAMLOracle(dims, 0.1, 0.01, case['inner_type'], dtype=torch.float64)
    o._eval_dtype = np.float64
    theta = case['theta'].astype(np.float64)

    def ev(th_np, want):
        t = torch.as_tensor(th_np, dtype=torch.float64).clone().requires_grad_(True)
        obj, _, okl = th.meta_objective(t, data, dims, 0.1, 'trpo', inner_type=case['inner_type'])
        (g,) = torch.autograd.grad(obj if want == 'grad' else okl, t)
        return g.numpy()
    assert rel_err(ev(theta, 'kl'), G[pre + 'kl_grad']) < 1e-6
    x = G[pre + 'hx_dir']
    eps = float(np.float32(1e-5))
    hx = (ev(theta + eps * x, 'kl') - ev(theta - eps * x, 'kl')) / (2 * eps)
    assert rel_err(hx, G[pre + 'hx']) < 1e-4


# ================================================================================================ GPU: CUDA path vs reference graph
def _cuda():
    import torch
    from promp_b200 import _lib
    _lib.require_cuda()
    return torch


def _product_algo(torch, case):
    from promp_b200.policies import MetaGaussianMLPPolicy
    from promp_b200.meta_algos import ProMP, TRPOMAML, VPGMAML
    H = tf_cases.HYPER
    M, S1 = case['M'], case['S'] - 1
    np.random.seed(1)
    policy = MetaGaussianMLPPolicy(name='meta-policy', obs_dim=case['Do'], action_dim=case['Da'], meta_batch_size=M,
                                   hidden_sizes=(case['hidden'], case['hidden']))
    policy.set_params(tf_cases.unflatten(case['theta'], case['Do'], case['Da'], case['hidden']))
    if case['algo'] == 'promp':
        algo = ProMP(policy=policy, inner_lr=H['inner_lr'], meta_batch_size=M, num_inner_grad_steps=S1,
                     learning_rate=H['learning_rate'], num_ppo_steps=H['num_ppo_steps'], clip_eps=H['clip_eps'],
                     target_inner_step=0.01, init_inner_kl_penalty=H['init_inner_kl_penalty'], adaptive_inner_kl_penalty=False)
    elif case['algo'] == 'trpo':
        algo = TRPOMAML(policy=policy, step_size=H['step_size'], inner_type=case['inner_type'], inner_lr=H['inner_lr'],
                        meta_batch_size=M, num_inner_grad_steps=S1, exploration=case.get('exploration', False))
    else:
        algo = VPGMAML(policy=policy, learning_rate=H['learning_rate'], inner_type=case['inner_type'], inner_lr=H['inner_lr'],
                       meta_batch_size=M, num_inner_grad_steps=S1, exploration=case.get('exploration', False))
    return policy, algo


@pytest.mark.gpu
@pytest.mark.parametrize('name', ALL)
def test_cuda_adapt_and_meta_gradient_match_reference_graph(golden_dir, name):
    """MAMLAlgo._adapt, the meta objective, its KLs and its second-order gradient on the device - fed the reference's own
    processed-sample dicts (INTEGRATION path, MAMLAlgo._phase_of) - against the unmodified reference graph (float64
    evaluation).  Bar: 1e-4 relative (gradient, gradient norm, adapted-parameter update)."""
    torch = _cuda()
    G = _gold(golden_dir)
    case = tf_cases.make_case(name)
    samples = tf_cases.reference_samples(case)
    policy, algo = _product_algo(torch, case)
    pre = name + '/f64/'
    keep = G[name + '/keep_tasks']
    th0 = case['theta'].astype(np.float64)
    policy.switch_to_pre_update()
    for s in range(case['S'] - 1):
        algo._adapt(samples[s])
        got = policy.theta_tasks.cpu().numpy().astype(np.float64)
        want = G[pre + 'adapt%d_tasks' % s]
        assert rel_err(got[keep] - th0, want - th0) < 1e-4, rel_err(got[keep] - th0, want - th0)
        delta = got - th0[None]
        np.testing.assert_allclose(np.sqrt((delta ** 2).sum(1)), G[pre + 'adapt%d_delta_norm' % s], rtol=1e-4)
        np.testing.assert_allclose(delta.sum(1), G[pre + 'adapt%d_delta_sum' % s], rtol=1e-3, atol=1e-5)
    phases = [algo._phase_of(s) for s in samples]
    if case['algo'] == 'trpo':
        g_got = algo.eval_gradient(policy.theta, phases, 'loss')
        loss, klv = algo.eval_scalars(policy.theta, phases)
        assert close(klv, G[pre + 'outer_kl'], 1e-3, 1e-7)
        gk = algo.eval_gradient(policy.theta, phases, 'kl')
        assert rel_err(gk, G[pre + 'kl_grad']) < 1e-4, rel_err(gk, G[pre + 'kl_grad'])
    else:
        res = algo._objective_pass(phases, want_grad=True)
        g_got = res['grad'].cpu().numpy()
        terms = algo.loss_terms(res).cpu().numpy()
        loss = terms[0]
        if case['algo'] == 'promp':
            S1 = case['S'] - 1
            np.testing.assert_allclose(terms[1:1 + S1], G[pre + 'inner_kl'], rtol=1e-3, atol=1e-7)
            assert close(terms[1 + S1], G[pre + 'outer_kl'], 1e-3, 1e-7)
    assert close(loss, G[pre + 'loss'], 1e-4, 2e-6), (loss, float(G[pre + 'loss']))
    err = rel_err(g_got, G[pre + 'grad'])
    assert err < 1e-4, err
    assert abs(np.linalg.norm(g_got) / np.linalg.norm(G[pre + 'grad']) - 1) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('name', PROMP + VPG)
def test_cuda_optimize_policy_matches_reference_graph(golden_dir, name):
    """ProMP.optimize_policy (5 epochs of TF1 Adam + stats pass) / VPGMAML.optimize_policy (1 epoch) on the device vs the
    unmodified reference's MAMLPPOOptimizer / MAMLFirstOrderOptimizer.  Adam normalises every coordinate's step to ~lr,
    so the compared quantity is the parameter UPDATE; the float32 reference's own distance to the float64 one is the
    noise scale."""
    torch = _cuda()
    G = _gold(golden_dir)
    case = tf_cases.make_case(name)
    samples = tf_cases.reference_samples(case)
    policy, algo = _product_algo(torch, case)
    key = 'theta_after_adam5' if case['algo'] == 'promp' else 'theta_after'
    want64, want32 = G[name + '/f64/' + key], G[name + '/f32/' + key]
    th0 = case['theta'].astype(np.float64)
    algo.optimize_policy(samples, log=False)
    got = policy.theta.cpu().numpy().astype(np.float64)
    ref_noise = rel_err(want32 - th0, want64 - th0)
    err = rel_err(got - th0, want64 - th0)
    assert err < max(2e-3, 3 * ref_noise), (err, ref_noise)
    np.testing.assert_allclose(got, want64, rtol=0, atol=2e-5)
    if case['algo'] == 'promp':
        ls = algo.last_stats
        pre = name + '/f64/'
        assert close(ls['loss_before'], G[pre + 'loss'], 1e-4, 2e-6)
        assert close(ls['loss_after'], G[pre + 'loss_after'], 2e-4, 2e-6)
        np.testing.assert_allclose(ls['inner_kls'], G[pre + 'inner_kl_after'], rtol=2e-3, atol=1e-7)
        assert close(ls['outer_kl'], G[pre + 'outer_kl_after'], 2e-3, 1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize('name', TRPO)
def test_cuda_trpo_matches_reference_graph(golden_dir, name):
    """TRPO-MAML on the device vs the unmodified reference: finite-difference Hx no noisier than the float32 reference's
    own (both measured against the float64 graph), CG direction, and the accepted step of optimize_policy."""
    torch = _cuda()
    G = _gold(golden_dir)
    case = tf_cases.make_case(name)
    samples = tf_cases.reference_samples(case)
    policy, algo = _product_algo(torch, case)
    phases = [algo._phase_of(s) for s in samples]
    p64, p32 = name + '/f64/', name + '/f32/'
    x = G[p64 + 'hx_dir'].astype(np.float32)
    flat = case['theta'].copy()
    hx = algo.optimizer.Hx(flat, phases, x)
    ref_noise = rel_err(G[p32 + 'hx'], G[p64 + 'hx'])
    err = rel_err(hx, G[p64 + 'hx'])
    assert err < max(2.0 * ref_noise, 2e-2), (err, ref_noise)
    algo.optimize_policy(samples, log=False)
    got = policy.theta.cpu().numpy().astype(np.float64)
    th0 = case['theta'].astype(np.float64)
    want64, want32 = G[p64 + 'theta_after'], G[p32 + 'theta_after']
    moved64, moved32 = np.abs(want64 - th0).max() > 0, np.abs(want32 - th0).max() > 0
    moved = np.abs(got - th0).max() > 0
    if moved64 == moved32:                     # the accept / reject decision is not at the mercy of float32 noise
        assert moved == moved64
    if moved and moved64:
        # the direction is a 10-iteration CG on a noisy float32 FD operator: compare step length (fixed by the KL
        # constraint) and direction cosine, with the float32 reference's own deviation as the scale
        step, step64 = got - th0, want64 - th0
        cos = float(step @ step64 / (np.linalg.norm(step) * np.linalg.norm(step64)))
        cos_ref = float((want32 - th0) @ step64 / (np.linalg.norm(want32 - th0) * np.linalg.norm(step64))) if moved32 else 1.0
        assert cos > min(0.98, 1 - 3 * (1 - cos_ref)), (cos, cos_ref)
        assert abs(np.linalg.norm(step) / np.linalg.norm(step64) - 1) < max(0.05, 3 * abs(np.linalg.norm(want32 - th0) / np.linalg.norm(step64) - 1))
        assert close(algo.last_stats['loss_after'], G[p64 + 'loss_after'], 0.05, 1e-5)


# ================================================================================================ end to end vs the reference Trainer
@pytest.mark.gpu
@pytest.mark.parametrize('rtype', ['dense', 'sparse'])
def test_trainer_matches_unmodified_reference_trainer(golden_dir, rtype):
    """promp_b200's Trainer / MetaSampler / MetaSampleProcessor / LinearFeatureBaseline / MetaGaussianMLPPolicy / ProMP
    against tests/golden/trainer_run.npz = 3 meta-iterations of the UNMODIFIED reference Trainer over the unmodified
    reference classes (BASELINE.json configs[0]; same numpy seed -> same tasks and reset states, same action noise).
    Compared: sampled goals (bit-exact), logged return statistics, losses / KLs, and theta after every iteration."""
    torch = _cuda()
    from promp_b200.envs import normalize, MetaPointEnvCorner
    from promp_b200.policies import MetaGaussianMLPPolicy
    from promp_b200.samplers import MetaSampler, MetaSampleProcessor
    from promp_b200.baselines import LinearFeatureBaseline
    from promp_b200.meta_algos import ProMP
    from promp_b200.meta_trainer import Trainer
    from promp_b200.utils import logger
    G = np.load(os.path.join(golden_dir, 'trainer_run.npz'))
    pre = rtype + '_'
    M, E, H, n_itr = 5, 4, 100, 3
    logger.set_quiet(True)
    env = normalize(MetaPointEnvCorner(reward_type=rtype))
    policy = MetaGaussianMLPPolicy(name='meta-policy', obs_dim=2, action_dim=2, meta_batch_size=M, hidden_sizes=(64, 64))
    policy.set_params(tf_cases.unflatten(G[pre + 'theta0'], 2, 2, 64))
    sampler = MetaSampler(env=env, policy=policy, rollouts_per_meta_task=E, meta_batch_size=M, max_path_length=H, parallel=False)
    proc = MetaSampleProcessor(baseline=LinearFeatureBaseline(), discount=0.99, gae_lambda=1, normalize_adv=True)
    algo = ProMP(policy=policy, inner_lr=0.1, meta_batch_size=M, num_inner_grad_steps=1, learning_rate=1e-3, num_ppo_steps=5,
                 clip_eps=0.3, target_inner_step=0.01, init_inner_kl_penalty=5e-4, adaptive_inner_kl_penalty=False)
    trainer = Trainer(algo=algo, env=env, sampler=sampler, sample_processor=proc, policy=policy, n_itr=n_itr, num_inner_grad_steps=1)
    noise = G[pre + 'noise']
    state = dict(itr=0, phase=0)
    orig = sampler.obtain_samples

    def obtain_with_reference_noise(*a, **k):
        sampler.inject(noise=noise[state['itr'], state['phase']])
        state['phase'] += 1
        return orig(*a, **k)
    sampler.obtain_samples = obtain_with_reference_noise
    keys = [str(k) for k in G[pre + 'log_keys']]
    want_logs = G[pre + 'log_vals']
    np.random.seed(1)
    th_prev = G[pre + 'theta0'].astype(np.float64)
    for itr in range(n_itr):
        state.update(itr=itr, phase=0)
        trainer.train_iteration(itr, log=True)
        kv = dict(logger.getkvs())
        logger.dumpkvs()
        goals = np.asarray(sampler.vec_env.tasks, dtype=np.float64)
        assert np.array_equal(goals, G[pre + 'goals'][itr])
        want = dict(zip(keys, want_logs[itr]))
        for k in keys:
            if 'Time' in k:
                continue
            assert k in kv, "logged key %s missing" % k
            tol = dict(rtol=2e-3, atol=2e-3) if 'Return' in k else dict(rtol=5e-3, atol=2e-6)
            np.testing.assert_allclose(float(kv[k]), want[k], err_msg=k, **tol)
        got = policy.theta.cpu().numpy().astype(np.float64)
        want_th = G[pre + 'thetas'][itr]
        upd, upd_want = got - th_prev, want_th - th_prev
        err = rel_err(upd, upd_want)
        assert err < 2e-2, (itr, err)          # Adam-normalised update of a closed-loop float32 rollout
        np.testing.assert_allclose(got, want_th, rtol=0, atol=3e-4)
        th_prev = want_th
        policy.set_params(tf_cases.unflatten(want_th.astype(np.float32), 2, 2, 64))   # re-sync so later iterations are compared like for like
    assert np.allclose(np.random.uniform(size=4), G[pre + 'rng_probe_after'])          # same numpy RNG consumption as the reference
